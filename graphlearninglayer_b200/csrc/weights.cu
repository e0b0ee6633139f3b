// K3 -- bandwidths, edge weights, degree and the linear system of the unlabeled block.
//   eps_i   : GLL.py:205 (auto: distance to the last kNN entry kappa(i)) / GLL.py:226 (fixed)
//   W_ij    : GLL.py:216 / 233   exp(-4 d^2 / eps_i / eps_j), evaluated in fp64, stored fp32
//   deg_i   : GLL.py:29          csgraph.laplacian degree (sum of the stored fp32 weights, so L 1 = 0 holds)
//   L_uu    : GLL.py:37,48       kept as diag = deg + tau and a compact CSR of the off-diagonal weights
//   rhs     : GLL.py:53          -L_ul Y = W_ul Y
// The dense n x n matrix C of GLL.py:209-213 is replaced by the map kappa[n].
#include <math.h>

#include "common.cuh"

namespace gll {
namespace {

__global__ void __launch_bounds__(256)
weights_kernel(const int* __restrict__ knn_idx, const float* __restrict__ knn_dist, const int* __restrict__ row_ptr,
               const int* __restrict__ col, const float* __restrict__ dist, const float* __restrict__ Y, int n, int k,
               int l, int lp, int k_lab, int eps_auto, float eps_fixed, float tau, float* __restrict__ eps,
               int* __restrict__ kappa, float* __restrict__ w, float* __restrict__ deg, int* __restrict__ uu_cnt,
               float* __restrict__ diag, float* __restrict__ rhs, float* __restrict__ ut, int* __restrict__ info) {
  const int i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const float ei_f = eps_auto ? __ldg(knn_dist + (size_t)i * k + (k - 1)) : eps_fixed;
  if (lane == 0) {
    eps[i] = ei_f;
    kappa[i] = eps_auto ? __ldg(knn_idx + (size_t)i * k + (k - 1)) : -1;
    if (ei_f < 1e-10f && info != nullptr) atomicOr(&info[GLL_INFO_STATUS], GLL_STATUS_EPS_TINY);
  }
  const double ei = (double)ei_f;
  const int e0 = row_ptr[i], e1 = row_ptr[i + 1];
  double dsum = 0.0;
  int uu = 0;
  for (int e = e0 + lane; e < e1; e += 32) {
    int j = col[e];
    double dd = (double)dist[e];
    double ej = eps_auto ? (double)__ldg(knn_dist + (size_t)j * k + (k - 1)) : (double)eps_fixed;
    float wv = (float)exp(-4.0 * dd * dd / ei / ej);
    w[e] = wv;
    dsum += (double)wv;
    uu += (j >= k_lab);
  }
  dsum = warp_sum(dsum);
  uu = warp_sum(uu);
  if (lane == 0) deg[i] = (float)dsum;
  if (i < k_lab) {
    // ut[:k_lab] = Y  (GLL.py:109)
    for (int c = lane; c < lp; c += 32) ut[(size_t)i * lp + c] = (c < l) ? Y[(size_t)i * l + c] : 0.f;
    return;
  }
  const int r = i - k_lab;
  if (lane == 0) {
    uu_cnt[r] = uu;
    diag[r] = (float)(dsum + (double)tau);
  }
  __syncwarp();
  // rhs_r = sum over labeled neighbours (sorted columns: they are the head of the row)
  const int e_lab_end = e1 - uu;
  for (int c = lane; c < lp; c += 32) {
    double acc = 0.0;
    if (c < l)
      for (int e = e0; e < e_lab_end; ++e) acc += (double)w[e] * (double)__ldg(Y + (size_t)col[e] * l + c);
    rhs[(size_t)r * lp + c] = (float)acc;
  }
}

__global__ void __launch_bounds__(256)
uu_fill_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, const float* __restrict__ w, int m, int k_lab,
               const int* __restrict__ uu_ptr, int* __restrict__ uu_col, float* __restrict__ uu_val, int* __restrict__ info) {
  const int r = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (r == 0 && lane == 0 && info != nullptr) info[GLL_INFO_NNZ_UU] = uu_ptr[m];
  if (r >= m) return;
  const int o0 = uu_ptr[r], cnt = uu_ptr[r + 1] - o0;
  const int src = row_ptr[k_lab + r + 1] - cnt;
  for (int t = lane; t < cnt; t += 32) {
    uu_col[o0 + t] = col[src + t] - k_lab;
    uu_val[o0 + t] = w[src + t];
  }
}

}  // namespace

size_t weights_ws_bytes(int n, int k) {
  (void)k;
  return align_up(sizeof(int) * (size_t)(n + 1), 256) + scan_ws_bytes(n + 1) + 1024;
}

int weights_run(const int* knn_idx, const float* knn_dist, const int* row_ptr, const int* col, const float* dist,
                const float* Y, int n, int k, int l, int k_lab, int eps_auto, float eps_fixed, float tau, float* eps,
                int* kappa, float* w, float* deg, int* uu_ptr, int* uu_col, float* uu_val, float* diag, float* rhs,
                float* ut, int* info, void* ws, size_t ws_bytes, cudaStream_t st) {
  GLL_REQUIRE(knn_idx && knn_dist && row_ptr && col && dist && eps && kappa && w && deg && uu_ptr && uu_col && uu_val &&
                  diag && rhs && ut && ws,
              "null pointer");
  GLL_REQUIRE(k_lab >= 0 && k_lab < n && l >= 1, "need 0 <= k_lab < n and l >= 1");
  GLL_REQUIRE(k_lab == 0 || Y != nullptr, "label matrix missing");
  if (ws_bytes < weights_ws_bytes(n, k)) {
    set_error("weights workspace too small: %zu < %zu", ws_bytes, weights_ws_bytes(n, k));
    return GLL_ERR_WORKSPACE;
  }
  const int m = n - k_lab, lp = padded_classes(l);
  Carver cv(ws, ws_bytes);
  int* uu_cnt = cv.take<int>(n + 1);
  void* scan_ws = cv.take<char>(scan_ws_bytes(n + 1));
  {
    GLL_PROF(KID_WEIGHTS, st);
    weights_kernel<<<ceil_div((long long)n * 32, 256), 256, 0, st>>>(knn_idx, knn_dist, row_ptr, col, dist, Y, n, k, l, lp,
                                                                   k_lab, eps_auto, eps_fixed, tau, eps, kappa, w, deg,
                                                                   uu_cnt, diag, rhs, ut, info);
  }
  GLL_LAUNCH_CHECK();
  int rc = exclusive_scan(uu_cnt, m, uu_ptr, scan_ws, st);
  if (rc) return rc;
  {
    GLL_PROF(KID_UU_FILL, st);
    uu_fill_kernel<<<ceil_div((long long)m * 32, 256), 256, 0, st>>>(row_ptr, col, w, m, k_lab, uu_ptr, uu_col, uu_val, info);
  }
  GLL_LAUNCH_CHECK();
  return GLL_OK;
}

}  // namespace gll

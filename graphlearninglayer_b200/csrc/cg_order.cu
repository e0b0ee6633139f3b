// Row schedule for the streaming CG kernel (cg.cu) on systems whose iterate does not fit in L2 (the 1M-node graph: 393 MB).
// Opt-in (GLL_B200_CG_ORDER=1; "force" for tests).
//
// The SpMV of an iteration gathers ~30 neighbour rows of 4*lp bytes per row; with the nodes in arbitrary order those
// gathers miss L2.  Nothing has to be relabelled to change that -- L2 caches scattered 400-byte rows as well as contiguous
// ones; what matters is WHICH rows the 148 CTAs work on at the same time.  So the rows are only SCHEDULED by cluster:
// key(i) = the class column that dominates the right-hand side of row i (forward: the labeled neighbours' classes,
// GLL.py:53; adjoint: the label of the batch row, GLL.py:93), rows with an all-zero right-hand side adopt the key of their
// first neighbour that has one; a stable 8-bit radix sort of (key, row) gives `order`, and the SpMV phase deals consecutive
// positions of `order` to the warps of the whole grid.  The schedule only changes which warp computes which row: every
// row's result is bit-identical, the dot products are summed in a different (still deterministic) order.
// Measured at n = 2^20, l = 100 (profiles/r01f_cg_streaming_experiments.md): the adjoint solve gains 8 %, the forward solve
// nothing -- the iteration is bound by the VOLUME crossing L2 -> SM (30 gathered rows per row), not by where it comes from.
// Kept because the multi-GPU row partition needs such an ordering to exchange halos instead of the whole iterate.
#include <cub/device/device_radix_sort.cuh>

#include <mutex>

#include "cg_common.cuh"

namespace gll {
namespace {

constexpr int ORDER_THREADS = 256;

// warp per row: key = argmax_c |rhs[i][c]| (first maximum), 255 if the row is all zero
__global__ void __launch_bounds__(ORDER_THREADS)
cg_row_key_kernel(const float* __restrict__ rhs, int m, int lp, unsigned char* __restrict__ key) {
  const int i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= m) return;
  float best = 0.f;
  int arg = 255;
  for (int c = lane; c < lp; c += 32) {
    const float v = fabsf(__ldg(rhs + (size_t)i * lp + c));
    if (v > best) {
      best = v;
      arg = c;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float ob = __shfl_xor_sync(FULL, best, o);
    const int oa = __shfl_xor_sync(FULL, arg, o);
    if (ob > best || (ob == best && oa < arg)) {
      best = ob;
      arg = oa;
    }
  }
  if (lane == 0) key[i] = (unsigned char)arg;
}

// thread per row: rows without a key adopt the first keyed neighbour's (one hop); also writes the identity permutation
__global__ void __launch_bounds__(ORDER_THREADS)
cg_row_key_fill_kernel(const int* __restrict__ ptr, const int* __restrict__ col, int m, const unsigned char* __restrict__ key,
                       unsigned char* __restrict__ key2, int* __restrict__ iota) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  unsigned char k = key[i];
  if (k == 255) {
    const int e1 = __ldg(ptr + i + 1);
    for (int e = __ldg(ptr + i); e < e1; ++e) {
      const unsigned char kj = key[__ldg(col + e)];
      if (kj != 255) {
        k = kj;
        break;
      }
    }
  }
  key2[i] = k;
  iota[i] = i;
}

// CUB's size query costs tens of microseconds of host time (device attribute and occupancy look-ups) and the workspace
// size is asked for several times per layer call: remember the last few sizes.
size_t sort_temp_bytes(int m) {
  static std::mutex mu;
  static int cached_m[8];
  static size_t cached_bytes[8];
  static int used = 0, next = 0;
  std::lock_guard<std::mutex> lock(mu);
  for (int t = 0; t < used; ++t)
    if (cached_m[t] == m) return cached_bytes[t];
  size_t bytes = 0;
  const cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned char*)nullptr, (unsigned char*)nullptr,
                                                        (const int*)nullptr, (int*)nullptr, m, 0, 8, (cudaStream_t)0);
  if (e != cudaSuccess || bytes == 0) {  // no device (CPU-only container) or a transient error: an upper bound, not cached
    (void)cudaGetLastError();
    return (size_t)m * 8 + ((size_t)1 << 20);
  }
  cached_m[next] = m;
  cached_bytes[next] = bytes;
  next = (next + 1) % 8;
  if (used < 8) ++used;
  return bytes;
}

}  // namespace

size_t cg_order_ws_bytes(int m) {
  return 3 * align_up((size_t)m, 256) + 2 * align_up(sizeof(int) * (size_t)m, 256) + align_up(sort_temp_bytes(m), 256) + 1024;
}

// order[m]: the row schedule.  ws must hold cg_order_ws_bytes(m) bytes beyond `order` itself.
int cg_row_order(const int* uu_ptr, const int* uu_col, const float* rhs, int m, int lp, int* order, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  if (ws_bytes < cg_order_ws_bytes(m)) {
    set_error("CG row-order workspace too small: %zu < %zu", ws_bytes, cg_order_ws_bytes(m));
    return GLL_ERR_WORKSPACE;
  }
  Carver cv(ws, ws_bytes);
  unsigned char* key = cv.take<unsigned char>(m);
  unsigned char* key2 = cv.take<unsigned char>(m);
  unsigned char* key_sorted = cv.take<unsigned char>(m);
  int* iota = cv.take<int>(m);
  size_t temp_bytes = sort_temp_bytes(m);
  void* temp = cv.take<char>(temp_bytes);
  {
    GLL_PROF(KID_CG_ORDER, st);
    cg_row_key_kernel<<<ceil_div((long long)m * 32, ORDER_THREADS), ORDER_THREADS, 0, st>>>(rhs, m, lp, key);
    cg_row_key_fill_kernel<<<ceil_div(m, ORDER_THREADS), ORDER_THREADS, 0, st>>>(uu_ptr, uu_col, m, key, key2, iota);
    GLL_LAUNCH_CHECK();
    GLL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, key2, key_sorted, iota, order, m, 0, 8, st));
  }
  return GLL_OK;
}

}  // namespace gll

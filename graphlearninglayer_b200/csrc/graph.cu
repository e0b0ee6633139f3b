// K2 -- symmetrised kNN graph as a CSR with sorted rows (replaces the scipy sequence at GLL.py:192-198:
// coo -> csr, D + D^T.*(D^T > D) - D.*(D^T > D), sparse.find).  Distances are bit-symmetric (knn.cu computes
// d_ij and d_ji identically), so the elementwise max reduces to the UNION of the directed kNN edges; exact
// zeros (self, duplicate points) are not edges because sparse.find drops them.
//
// count -> exclusive scan -> fill (forward edges at deterministic slots, reverse edges through an integer cursor)
// -> per-row rank sort by column, which removes the only order nondeterminism.  Also hosts the prefix-sum.
#include "common.cuh"

namespace gll {
namespace {

// ------------------------------------------------------------------------------------------------ scan
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /* [32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int s = (lane < nw) ? smem[lane] : 0;
    int sinc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(FULL, sinc, o);
      if (lane >= o) sinc += t;
    }
    smem[lane] = sinc - s;  // exclusive warp offsets
    if (lane == 31) smem[32] = sinc;
  }
  __syncthreads();
  int res = inc - v + smem[warp];
  *total = smem[32];
  __syncthreads();
  return res;
}

// one block walks the whole array with a running carry (n small) -- a single launch
__global__ void __launch_bounds__(1024) scan_single_kernel(const int* __restrict__ in, int n, int* __restrict__ out) {
  __shared__ int sm[33];
  int carry = 0;
  for (int base = 0; base < n; base += 1024 * 4) {
    int i0 = base + threadIdx.x * 4;
    int v[4], s = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      v[t] = (i0 + t < n) ? in[i0 + t] : 0;
      s += v[t];
    }
    int tot;
    int ex = block_exclusive_scan(s, &tot, sm) + carry;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (i0 + t < n) out[i0 + t] = ex;
      ex += v[t];
    }
    carry += tot;
  }
  if (threadIdx.x == 0) out[n] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(const int* __restrict__ in, int n, int* __restrict__ out,
                                                                 int* __restrict__ tile_sums) {
  __shared__ int sm[33];
  int i0 = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int t = 0; t < SCAN_ITEMS; ++t) {
    v[t] = (i0 + t < n) ? in[i0 + t] : 0;
    s += v[t];
  }
  int tot;
  int ex = block_exclusive_scan(s, &tot, sm);
#pragma unroll
  for (int t = 0; t < SCAN_ITEMS; ++t) {
    if (i0 + t < n) out[i0 + t] = ex;
    ex += v[t];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(int* __restrict__ out, int n,
                                                                const int* __restrict__ tile_offsets) {
  int off = tile_offsets[blockIdx.x];
  int i0 = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int t = 0; t < SCAN_ITEMS; ++t)
    if (i0 + t < n) out[i0 + t] += off;
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = tile_offsets[gridDim.x];
}

}  // namespace

size_t scan_ws_bytes(int n) { return align_up(sizeof(int) * (size_t)(2 * (ceil_div(n, SCAN_TILE) + 2)), 256); }

int exclusive_scan(const int* counts, int n, int* out, void* scratch, cudaStream_t st) {
  if (n <= 32768) {
    {
      GLL_PROF(KID_SCAN, st);
      scan_single_kernel<<<1, 1024, 0, st>>>(counts, n, out);
    }
    GLL_LAUNCH_CHECK();
    return GLL_OK;
  }
  int tiles = ceil_div(n, SCAN_TILE);
  int* sums = (int*)scratch;
  int* offs = sums + tiles + 1;
  {
    GLL_PROF(KID_SCAN, st);
    scan_tile_kernel<<<tiles, SCAN_THREADS, 0, st>>>(counts, n, out, sums);
  }
  GLL_LAUNCH_CHECK();
  {
    GLL_PROF(KID_SCAN, st);
    scan_single_kernel<<<1, 1024, 0, st>>>(sums, tiles, offs);
  }
  GLL_LAUNCH_CHECK();
  {
    GLL_PROF(KID_SCAN, st);
    scan_add_kernel<<<tiles, SCAN_THREADS, 0, st>>>(out, n, offs);
  }
  GLL_LAUNCH_CHECK();
  return GLL_OK;
}

namespace {

// ------------------------------------------------------------------------------------------------ graph
__device__ __forceinline__ bool is_edge(int i, int j, float dd, int n) { return dd > 0.f && j != i && j >= 0 && j < n; }

// Does row j list i as a valid (non-zero) neighbour?
__device__ __forceinline__ bool lists(const int* __restrict__ knn_idx, const float* __restrict__ knn_dist, int n, int k,
                                      int j, int i) {
  // distances are bit-symmetric (knn.cu), so an entry i in row j carries the same non-zero distance as (i, j): only the
  // indices need to be scanned
  (void)knn_dist;
  (void)n;
  bool found = false;
  const int* ri = knn_idx + (size_t)j * k;
  for (int t = 0; t < k; ++t) found |= (__ldg(ri + t) == i);
  return found;
}

// flag per kNN entry: 0 not an edge, 1 edge listed by both ends, 2 edge listed only by i (the reverse edge j->i must be
// created).  The (expensive) membership scan is done once here and reused by the fill kernel.
__global__ void graph_count_kernel(const int* __restrict__ knn_idx, const float* __restrict__ knn_dist, int n, int k,
                                   int* __restrict__ len, unsigned char* __restrict__ flag) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * k) return;
  int i = (int)(t / k);
  int j = knn_idx[t];
  float dd = knn_dist[t];
  if (!is_edge(i, j, dd, n)) {
    flag[t] = 0;
    return;
  }
  atomicAdd(&len[i], 1);
  const bool mutual = lists(knn_idx, knn_dist, n, k, j, i);
  if (!mutual) atomicAdd(&len[j], 1);
  flag[t] = mutual ? 1 : 2;
}

// forward edges of row i fill the head of the row in kNN order (deterministic slots); reverse edges fill the tail through an
// integer cursor (their order is fixed afterwards by the per-row sort)
__global__ void graph_fill_kernel(const int* __restrict__ knn_idx, const float* __restrict__ knn_dist, int n, int k,
                                  const int* __restrict__ row_ptr, const unsigned char* __restrict__ flag,
                                  int* __restrict__ cursor, int* __restrict__ col_tmp, float* __restrict__ dist_tmp) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * k) return;
  const unsigned char f = flag[t];
  if (f == 0) return;
  int i = (int)(t / k), s = (int)(t % k);
  int j = knn_idx[t];
  float dd = knn_dist[t];
  int before = 0;
  for (int u = 0; u < s; ++u) before += (flag[(size_t)i * k + u] != 0);
  int p = row_ptr[i] + before;
  col_tmp[p] = j;
  dist_tmp[p] = dd;
  if (f == 2) {
    int q = row_ptr[j + 1] - 1 - atomicAdd(&cursor[j], 1);
    col_tmp[q] = i;
    dist_tmp[q] = dd;
  }
}

// warp per row: out-of-place rank sort by column (columns are unique within a row)
__global__ void graph_sort_rows_kernel(const int* __restrict__ row_ptr, int n, const int* __restrict__ col_tmp,
                                       const float* __restrict__ dist_tmp, int* __restrict__ col, float* __restrict__ dist,
                                       int* __restrict__ info) {
  int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0 && info != nullptr) info[GLL_INFO_NNZ] = row_ptr[n];
  if (warp >= n) return;
  int e0 = row_ptr[warp], len = row_ptr[warp + 1] - e0;
  for (int e = lane; e < len; e += 32) {
    int c = col_tmp[e0 + e];
    int rank = 0;
    for (int f = 0; f < len; ++f) rank += (col_tmp[e0 + f] < c);
    col[e0 + rank] = c;
    dist[e0 + rank] = dist_tmp[e0 + e];
  }
}

}  // namespace

size_t graph_ws_bytes(int n, int k) {
  size_t emax = gll_max_edges(n, k);
  return align_up(sizeof(int) * (size_t)(n + 1), 256) * 2 + align_up(sizeof(int) * emax, 256) +
         align_up(sizeof(float) * emax, 256) + scan_ws_bytes(n + 1) + align_up((size_t)n * k, 256) + 1024;
}

int graph_run(const int* knn_idx, const float* knn_dist, int n, int k, int* row_ptr, int* col, float* dist, int* info,
              void* ws, size_t ws_bytes, cudaStream_t st) {
  GLL_REQUIRE(knn_idx && knn_dist && row_ptr && col && dist && ws, "null pointer");
  GLL_REQUIRE(n >= 1 && k >= 2, "bad sizes");
  if (ws_bytes < graph_ws_bytes(n, k)) {
    set_error("graph workspace too small: %zu < %zu", ws_bytes, graph_ws_bytes(n, k));
    return GLL_ERR_WORKSPACE;
  }
  size_t emax = gll_max_edges(n, k);
  Carver cv(ws, ws_bytes);
  int* len = cv.take<int>(n + 1);
  int* cursor = cv.take<int>(n + 1);
  int* col_tmp = cv.take<int>(emax);
  float* dist_tmp = cv.take<float>(emax);
  void* scan_ws = cv.take<char>(scan_ws_bytes(n + 1));
  unsigned char* flag = cv.take<unsigned char>((size_t)n * k);
  // len and cursor are adjacent 256-aligned blocks: clear both with one memset
  GLL_CUDA_CHECK(cudaMemsetAsync(len, 0, (size_t)((char*)cursor - (char*)len) + sizeof(int) * (size_t)(n + 1), st));
  long long total = (long long)n * k;
  int blocks = ceil_div(total, 256);
  {
    GLL_PROF(KID_GRAPH_COUNT, st);
    graph_count_kernel<<<blocks, 256, 0, st>>>(knn_idx, knn_dist, n, k, len, flag);
  }
  GLL_LAUNCH_CHECK();
  int rc = exclusive_scan(len, n, row_ptr, scan_ws, st);
  if (rc) return rc;
  {
    GLL_PROF(KID_GRAPH_FILL, st);
    graph_fill_kernel<<<blocks, 256, 0, st>>>(knn_idx, knn_dist, n, k, row_ptr, flag, cursor, col_tmp, dist_tmp);
  }
  GLL_LAUNCH_CHECK();
  {
    GLL_PROF(KID_GRAPH_SORT, st);
    graph_sort_rows_kernel<<<ceil_div((long long)n * 32, 256), 256, 0, st>>>(row_ptr, n, col_tmp, dist_tmp, col, dist, info);
  }
  GLL_LAUNCH_CHECK();
  return GLL_OK;
}

}  // namespace gll

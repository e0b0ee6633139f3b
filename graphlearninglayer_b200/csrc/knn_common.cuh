// Pieces shared by the SIMT (knn.cu) and tensor-core (knn_tc.cu) kNN candidate kernels.
#pragma once
#include "common.cuh"

namespace gll {

typedef unsigned long long u64;
constexpr int KC = 32;               // candidates kept per row (one per lane)
constexpr int KNN_MAX_SPLITS = 32;   // candidate lists per row: two per CTA that touches the row tile
constexpr u64 KEY_INF = ~0ull;

// key = (order-preserving bits of the approximate squared distance, column index): u64 compare == (dist, idx) compare
__device__ __forceinline__ u64 make_key(float dist, int j) {
  return ((u64)float_to_ordered(dist) << 32) | (uint32_t)j;
}
__device__ __forceinline__ float key_dist(u64 k) { return ordered_to_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ int key_idx(u64 k) { return (int)(uint32_t)k; }

// Insert x into an ascending list of 32 keys held one per lane (keys are unique). No-op if x is larger than all.
__device__ __forceinline__ void list_insert(u64& mine, u64 x, int lane) {
  int pos = __popc(__ballot_sync(FULL, mine < x));
  u64 prev = __shfl_up_sync(FULL, mine, 1);
  if (lane > pos)
    mine = prev;
  else if (lane == pos)
    mine = x;
}

// compare-exchange with the lane at xor distance j: keep the smaller key if keep_min
__device__ __forceinline__ void key_cmpx(u64& k, int j, bool keep_min) {
  const u64 o = __shfl_xor_sync(FULL, k, j);
  const bool less = k < o;
  k = (less == keep_min) ? k : o;
}
// ascending bitonic sort of 32 keys, one per lane
__device__ __forceinline__ void key_sort32(u64& k, int lane) {
#pragma unroll
  for (int w = 2; w <= 32; w <<= 1) {
#pragma unroll
    for (int j = w >> 1; j >= 1; j >>= 1) {
      const bool up = (lane & w) == 0;  // w == 32: every lane ascending
      key_cmpx(k, j, ((lane & j) == 0) == up);
    }
  }
}
// Merge one unsorted 32-set (one key per lane, KEY_INF = empty) into the sorted list `mine`.  Only keys below the list's
// current last entry can enter: a handful are inserted one by one (one ballot finds them); many -- the first lists of a
// row -- go through a sorting network: sort the set, take min(mine[i], set[31 - i]) (the 32 smallest of the union, a
// bitonic sequence) and merge it back into ascending order.  The serial form alone was 40 % of the re-rank kernel's
// instructions.
__device__ __forceinline__ void list_merge_set(u64& mine, u64 c, int lane) {
  unsigned m = __ballot_sync(FULL, c < __shfl_sync(FULL, mine, KC - 1));
  if (m == 0) return;
  if (__popc(m) > 6) {
    key_sort32(c, lane);
    const u64 r = __shfl_sync(FULL, c, 31 - lane);
    mine = (r < mine) ? r : mine;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) key_cmpx(mine, j, (lane & j) == 0);
    return;
  }
  while (m) {
    const int t = __ffs(m) - 1;
    m &= m - 1;
    const u64 x = __shfl_sync(FULL, c, t);
    if (x < __shfl_sync(FULL, mine, KC - 1)) list_insert(mine, x, lane);  // the last entry may have tightened meanwhile
  }
}

// Tensor-core candidate generation (knn_tc.cu).  plan.ok == 0: shape or configuration not handled, use the SIMT path.
struct TcPlan {
  int ok, d_pad, kblocks, row_tiles, col_tiles, grid, max_splits, rt0, aligned, rstep;  // row_tiles counts row GROUPS of rstep tiles
  int ct0, col_begin;  // column range [col_begin, n): first column tile; columns before col_begin inside it are masked
  int passes;       // MMA passes over the fp16 operands (rows of X scaled by a power of two): 1 = hi.hi, 2 = (hi + lo).hi
  long long units;
  size_t ws_bytes;  // fp16 hi / lo copies of X
};

TcPlan knn_tc_plan(int n, int d, int row_begin, int row_end, int col_begin = 0);
size_t knn_tc_ws_upper(int n, int d);
int knn_tc_candidates(const float* X, const float* sq, const float* rscale, const unsigned* small, int n, int d, int row_end, const TcPlan& plan,
                      void* tc_ws, u64* cand, const u64* excl, unsigned* thr_g, cudaStream_t st);
float knn_tc_err_coef(int d, int passes);
// verification: raw accumulator of one (row tile, column tile) unit, acc_out[128][256] (knn_gram_tile_debug_kernel)
int knn_tc_debug_tile(const TcPlan& plan, int n, void* tc_ws, int rt, int ct, float* acc_out, cudaStream_t st);

// How the candidate lists of a row are laid out in cand[n][stride][KC]: uniform (SIMT: every row has `stride` lists)
// or the tensor-core work split (row tile rt was touched by CTAs b0(rt)..b1(rt); slot = b - b0).
struct CandLayout {
  int stride;      // lists allocated per row
  int tc;          // 0: every row has `stride` sorted lists; 1: tensor-core work split; 2: one unsorted set per row
  int row_begin;   // first row of this call's row range (row tiles are counted from it)
  int row_tile, col_tiles, grid;
  long long units;
};

int knn_finish(const float* X, const float* sq, const unsigned* sqmax_bits, int n, int d, int k, int row_begin, int row_end,
               CandLayout lay,
               const u64* cand, float err_coef, int* knn_idx, float* knn_dist, int* flag_count, int* flag_rows,
               int* info, void* fb_scratch, cudaStream_t st);

}  // namespace gll

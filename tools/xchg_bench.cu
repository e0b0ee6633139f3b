// Microbenchmark of grid-wide exchange protocols between co-resident CTAs on one GPU (B200, sm_100a).
// It answers the questions behind the on-chip CG kernel's two exchanges per iteration (csrc/cg_resident.cu):
// how long is one L2 hop between two SMs, what do an all-to-all "everybody published" exchange and an all-reduce of
// 3*l scalars cost with 148 participants, and what do thread-block clusters (hardware barrier + DSMEM) buy.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/xchg_bench tools/xchg_bench.cu
//   tools/xchg_bench            (prints one line per protocol: ns per exchange)
//
// Every test runs REPS exchanges inside one cooperative launch and is timed with CUDA events.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__);     \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

constexpr int THREADS = 1024;
constexpr int NCOL = 30;  // 3 dot products x 10 classes
constexpr int MAXG = 160;

__device__ __forceinline__ unsigned ld_rlx(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_rlx(unsigned* p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_rel(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_rlx64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_rlx64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add_rel(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__device__ __forceinline__ void cluster_sync_() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank_() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_id_() {
  unsigned r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned map_to_rank(const void* smem_ptr, unsigned rank) {
  unsigned a = (unsigned)__cvta_generic_to_shared(smem_ptr), out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(a), "r"(rank));
  return out;
}
__device__ __forceinline__ void st_cluster_f32(unsigned addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

__device__ __forceinline__ float4 ldcg_v4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

struct Bufs {
  unsigned* flags;              // [MAXG][MAXG]
  unsigned long long* words;    // [NCOL][MAXG]
  unsigned long long* results;  // [MAXG][NCOL]
  unsigned* counter;            // [64]
  float* sink;
  int reps;
  int mode;
  int fence;  // 0 none, 1 threadfence both sides
  int sleep_ns;
  float* u;    // gather test: [GM][stride] floats
  int stride;  // floats per row of u (12 = packed 48-byte rows, 16 = 64-byte rows)
  int seg;     // edges per thread
};
constexpr int GM = 14336;      // rows of the gathered iterate (C4)
constexpr int GE = 2560;       // gathers per CTA and pass (C4: 2545 on average)
constexpr int GQ = 3;          // class quads per row (lp = 12)

__shared__ __align__(16) unsigned char g_smem[MAXG * NCOL * 8];  // 38 KB scratch shared by the tests (one runs at a time)

__device__ __forceinline__ void backoff(int ns) {
  if (ns > 0) __nanosleep(ns);
}

// ---- T1: ping-pong between CTA 0 and CTA G-1 ----
__device__ void t_pingpong(const Bufs& B) {
  const int G = gridDim.x, b = blockIdx.x;
  if (threadIdx.x != 0) return;
  unsigned* mine = B.flags + b * MAXG;
  if (b == 0) {
    for (int k = 1; k <= B.reps; ++k) {
      st_rlx(B.flags + (G - 1) * MAXG, k);
      while (ld_rlx(mine) < (unsigned)k) backoff(B.sleep_ns);
    }
  } else if (b == G - 1) {
    for (int k = 1; k <= B.reps; ++k) {
      while (ld_rlx(mine) < (unsigned)k) backoff(B.sleep_ns);
      st_rlx(B.flags, k);
    }
  }
}

// ---- T2: all-to-all "published" flags, per-destination mailboxes (the CG kernel's E1) ----
__device__ void t_flags_mailbox(const Bufs& B) {
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = 1; k <= B.reps; ++k) {
    __syncthreads();
    if (warp == 0) {
      if (B.fence) __threadfence();
      for (int d = lane; d < G; d += 32) st_rlx(B.flags + d * MAXG + b, k);
    }
    if (warp == 1) {
      const unsigned* mine = B.flags + b * MAXG;
      for (int s = lane; s < G; s += 32)
        while (ld_rlx(mine + s) < (unsigned)k) backoff(B.sleep_ns);
      if (B.fence) __threadfence();
    }
    __syncthreads();
  }
}

// ---- T3: shared flag array: CTA b writes flags[b]; everybody polls all G flags (shared lines) ----
__device__ void t_flags_shared(const Bufs& B) {
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = 1; k <= B.reps; ++k) {
    __syncthreads();
    if (tid == 0) {
      if (B.fence) __threadfence();
      st_rlx(B.flags + b, k);
    }
    if (warp == 1) {
      for (int s = lane; s < G; s += 32)
        while (ld_rlx(B.flags + s) < (unsigned)k) backoff(B.sleep_ns);
      if (B.fence) __threadfence();
    }
    __syncthreads();
  }
}

// ---- T4: one atomic counter (classic grid barrier) ----
__device__ void t_counter(const Bufs& B) {
  const int G = gridDim.x;
  for (int k = 1; k <= B.reps; ++k) {
    __syncthreads();
    if (threadIdx.x == 0) {
      red_add_rel(B.counter, 1u);
      while (ld_acq(B.counter) < (unsigned)(k * G)) backoff(B.sleep_ns);
    }
    __syncthreads();
  }
}

// ---- T5: all-reduce of NCOL scalars, two hops through owner CTAs (the CG kernel's E2) ----
__device__ float t_reduce_twohop(const Bufs& B) {
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ float red[NCOL];
  float acc = 0.f;
  for (int k = 1; k <= B.reps; ++k) {
    __syncthreads();
    if (tid < NCOL) {
      const unsigned long long w = ((unsigned long long)k << 32) | __float_as_uint(1.0f + tid);
      st_rlx64(B.words + (size_t)tid * MAXG + b, w);
    }
    for (int kk = warp; b + kk * G < NCOL; kk += THREADS / 32) {
      const int cr = b + kk * G;
      const unsigned long long* wb = B.words + (size_t)cr * MAXG;
      double t = 0.0;
      for (int s = lane; s < G; s += 32) {
        unsigned long long w;
        do {
          w = ld_rlx64(wb + s);
          if ((unsigned)(w >> 32) != (unsigned)k) backoff(B.sleep_ns);
        } while ((unsigned)(w >> 32) != (unsigned)k);
        t += (double)__uint_as_float((unsigned)w);
      }
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      const unsigned long long out = ((unsigned long long)k << 32) | __float_as_uint((float)t);
      for (int d = lane; d < G; d += 32) st_rlx64(B.results + (size_t)d * NCOL + cr, out);
    }
    if (warp == THREADS / 32 - 1) {
      const unsigned long long* mine = B.results + (size_t)b * NCOL;
      for (int cr = lane; cr < NCOL; cr += 32) {
        unsigned long long w;
        do {
          w = ld_rlx64(mine + cr);
          if ((unsigned)(w >> 32) != (unsigned)k) backoff(B.sleep_ns);
        } while ((unsigned)(w >> 32) != (unsigned)k);
        red[cr] = __uint_as_float((unsigned)w);
      }
    }
    __syncthreads();
    acc += red[tid % NCOL];
  }
  return acc;
}

// ---- T6: all-reduce, ONE hop: everybody writes its row of tagged words {fp32, epoch} (double buffered by the parity of
// the exchange: nobody can be two exchanges ahead of anybody else), everybody polls ALL G rows with coalesced loads and
// adds them itself in a fixed order ----
__device__ float t_reduce_onehop(const Bufs& B) {
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* stage = reinterpret_cast<float*>(g_smem);
  __shared__ float red[NCOL];
  float acc = 0.f;
  for (int k = 1; k <= B.reps; ++k) {
    unsigned long long* buf = B.words + (size_t)(k & 1) * MAXG * NCOL;
    __syncthreads();
    if (tid < NCOL) st_rlx64(buf + (size_t)b * NCOL + tid, ((unsigned long long)k << 32) | __float_as_uint(1.0f + tid));
    for (int idx = tid; idx < G * NCOL; idx += THREADS) {
      unsigned long long w;
      do {
        w = ld_rlx64(buf + idx);
        if ((unsigned)(w >> 32) != (unsigned)k) backoff(B.sleep_ns);
      } while ((unsigned)(w >> 32) != (unsigned)k);
      stage[idx] = __uint_as_float((unsigned)w);
    }
    __syncthreads();
    if (warp < NCOL) {  // warp c sums column c over all CTAs (fixed order)
      float t = 0.f;
      for (int s = lane; s < G; s += 32) t += stage[s * NCOL + warp];
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) red[warp] = t;
    }
    __syncthreads();
    acc += red[tid % NCOL];
  }
  return acc;
}

// ---- T6b: the counter barrier followed by a coalesced read of all G rows of plain doubles (what the CG kernel does) ----
__device__ float t_reduce_barrier_read(const Bufs& B) {
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* stage = reinterpret_cast<double*>(g_smem);
  __shared__ float red[NCOL];
  double* rows = reinterpret_cast<double*>(B.words);
  float acc = 0.f;
  for (int k = 1; k <= B.reps; ++k) {
    __syncthreads();
    if (tid < NCOL) rows[(size_t)b * NCOL + tid] = 1.0 + tid + k;
    __syncthreads();
    if (tid == 0) {
      red_add_rel(B.counter, 1u);
      while (ld_acq(B.counter) < (unsigned)(k * G)) {
      }
    }
    __syncthreads();
    for (int idx = tid; idx < G * NCOL; idx += THREADS) stage[idx] = __ldcg(rows + idx);
    __syncthreads();
    if (warp < NCOL) {
      float t = 0.f;
      for (int s = lane; s < G; s += 32) t += (float)stage[s * NCOL + warp];
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) red[warp] = t;
    }
    __syncthreads();
    acc += red[tid % NCOL];
    // the next exchange overwrites rows[]: nobody may still be reading -> in the CG kernel barrier (B) lies in between
    __syncthreads();
    if (tid == 0) {
      red_add_rel(B.counter + 32, 1u);
      while (ld_acq(B.counter + 32) < (unsigned)(k * G)) {
      }
    }
  }
  return acc;
}

// ---- T6c: ONE hop, doubles as two tagged 64-bit words {hi32|epoch}, {lo32|epoch}; column-major [col][cta][2]; one warp per
// column polls its 148 entries with all loads in flight (16 bytes per lane and load), then one butterfly ----
__device__ __forceinline__ void ld_rlx_v2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_rlx_v2(unsigned long long* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ float t_reduce_onehop_cols(const Bufs& B) {
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double red[NCOL];
  float acc = 0.f;
  for (int k = 1; k <= B.reps; ++k) {
    unsigned long long* buf = B.words + (size_t)(k & 1) * MAXG * NCOL * 2;
    __syncthreads();
    if (warp < NCOL && lane == 0) {  // in the CG kernel: lane 0 of the warp that computed this column's partial
      const unsigned long long bits = (unsigned long long)__double_as_longlong(1.0 + warp + 1e-9 * b);
      st_rlx_v2(buf + ((size_t)warp * MAXG + b) * 2, ((unsigned long long)k << 32) | (bits >> 32), ((unsigned long long)k << 32) | (bits & 0xffffffffull));
    }
    if (warp < NCOL) {
      const unsigned long long* src = buf + (size_t)warp * MAXG * 2;
      unsigned long long hi[5], lo[5];
      bool done[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) done[j] = lane + 32 * j >= G;
      bool all;
      do {
        all = true;
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (!done[j]) {
            ld_rlx_v2(src + (size_t)(lane + 32 * j) * 2, hi[j], lo[j]);
            done[j] = (unsigned)(hi[j] >> 32) == (unsigned)k && (unsigned)(lo[j] >> 32) == (unsigned)k;
            all &= done[j];
          }
      } while (!__all_sync(0xffffffffu, all));
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (lane + 32 * j < G) t += __longlong_as_double((long long)((hi[j] << 32) | (lo[j] & 0xffffffffull)));
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) red[warp] = t;
    }
    __syncthreads();
    acc += (float)red[tid % NCOL];
  }
  return acc;
}

// ---- T6d: what the CG kernel does now: plain doubles column-major, counter barrier, one warp per column reads with all
// loads in flight, butterfly; plus the second barrier of the iteration (the buffer is reused) ----
__device__ float t_reduce_barrier_cols(const Bufs& B) {
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double red[NCOL];
  double* cols = reinterpret_cast<double*>(B.words);
  float acc = 0.f;
  for (int k = 1; k <= B.reps; ++k) {
    if (warp < NCOL && lane == 0) cols[(size_t)warp * MAXG + b] = 1.0 + warp + k;
    __syncthreads();
    if (tid == 0) {
      red_add_rel(B.counter, 1u);
      while (ld_acq(B.counter) < (unsigned)((2 * k - 1) * G)) {
      }
    }
    __syncthreads();
    if (warp < NCOL) {
      const double* src = cols + (size_t)warp * MAXG;
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (lane + 32 * j < G) t += __ldcg(src + lane + 32 * j);
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) red[warp] = t;
    }
    __syncthreads();
    acc += (float)red[tid % NCOL];
    if (tid == 0) {
      red_add_rel(B.counter, 1u);
      while (ld_acq(B.counter) < (unsigned)(2 * k * G)) {
      }
    }
    __syncthreads();
  }
  return acc;
}

// ---- T7: cluster-hierarchical flags: cluster barrier, leaders exchange mailbox flags, cluster barrier ----
__device__ void t_flags_cluster(const Bufs& B) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned crank = cluster_rank_(), cid = cluster_id_();
  unsigned ncl;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(ncl));
  for (int k = 1; k <= B.reps; ++k) {
    if (B.fence && tid == 0) __threadfence();
    cluster_sync_();
    if (crank == 0) {
      if (warp == 0)
        for (int d = lane; d < (int)ncl; d += 32) st_rlx(B.flags + d * MAXG + cid, k);
      if (warp == 1) {
        const unsigned* mine = B.flags + cid * MAXG;
        for (int s = lane; s < (int)ncl; s += 32)
          while (ld_rlx(mine + s) < (unsigned)k) backoff(B.sleep_ns);
        if (B.fence) __threadfence();
      }
    }
    cluster_sync_();
  }
}

// ---- T8: cluster-hierarchical all-reduce: DSMEM partials -> leader, leaders one hop through L2, DSMEM broadcast ----
__device__ float t_reduce_cluster(const Bufs& B) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned crank = cluster_rank_(), cid = cluster_id_();
  unsigned ncl, csz;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(ncl));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csz));
  __shared__ float part[16][NCOL];  // leader: partials of the cluster's CTAs
  __shared__ float red[NCOL];       // every CTA: the result
  float acc = 0.f;
  for (int k = 1; k <= B.reps; ++k) {
    if (tid < NCOL) st_cluster_f32(map_to_rank(&part[crank][tid], 0), 1.0f + tid);
    cluster_sync_();
    if (crank == 0) {
      if (tid < NCOL) {
        float t = 0.f;
        for (unsigned c = 0; c < csz; ++c) t += part[c][tid];
        const unsigned long long w = ((unsigned long long)k << 32) | __float_as_uint(t);
        st_rlx64(B.results + (size_t)cid * NCOL + tid, w);
      }
      if (warp >= 1 && warp <= NCOL) {  // warp c+1 sums column c over the cluster leaders
        const int c = warp - 1;
        double t = 0.0;
        for (int s = lane; s < (int)ncl; s += 32) {
          unsigned long long w;
          do {
            w = ld_rlx64(B.results + (size_t)s * NCOL + c);
            if ((unsigned)(w >> 32) != (unsigned)k) backoff(B.sleep_ns);
          } while ((unsigned)(w >> 32) != (unsigned)k);
          t += (double)__uint_as_float((unsigned)w);
        }
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane < (int)csz) st_cluster_f32(map_to_rank(&red[c], lane), (float)t);
      }
    }
    cluster_sync_();
    acc += red[tid % NCOL];
  }
  return acc;
}

// ---- T9: cooperative-groups grid.sync() for reference ----
__device__ void t_gridsync(const Bufs& B) {
  cg::grid_group g = cg::this_grid();
  for (int k = 1; k <= B.reps; ++k) g.sync();
}

// ---- T10: cluster barrier alone ----
__device__ void t_clusterbar(const Bufs& B) {
  for (int k = 1; k <= B.reps; ++k) cluster_sync_();
}

// ---- T11: the SpMV gather of the CG kernel in isolation: thread = (segment of `seg` edges, class quad), all loads of a
// thread independent; no exchange between passes (one __syncthreads) ----
__device__ float t_gather(const Bufs& B) {
  int* col = reinterpret_cast<int*>(g_smem);                    // [GE]
  float4* out = reinterpret_cast<float4*>(g_smem + GE * 4);      // [THREADS]
  const int tid = threadIdx.x;
  for (int e = tid; e < GE; e += THREADS) {
    unsigned h = (unsigned)(e * 2654435761u) ^ (blockIdx.x * 40503u + 12345u);
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    col[e] = (int)(h % GM);
  }
  __syncthreads();
  const int nseg = GE / B.seg, tasks = nseg * GQ;
  float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 1; k <= B.reps; ++k) {
    for (int t = tid; t < tasks; t += THREADS) {
      const int sg = t / GQ, q = t - sg * GQ;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (B.seg == 8) {
        float4 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = ldcg_v4(B.u + (size_t)col[sg * 8 + e] * B.stride + 4 * q);
#pragma unroll
        for (int e = 0; e < 8; ++e) { a.x += v[e].x; a.y += v[e].y; a.z += v[e].z; a.w += v[e].w; }
      } else if (B.seg == 4) {
        float4 v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = ldcg_v4(B.u + (size_t)col[sg * 4 + e] * B.stride + 4 * q);
#pragma unroll
        for (int e = 0; e < 4; ++e) { a.x += v[e].x; a.y += v[e].y; a.z += v[e].z; a.w += v[e].w; }
      } else {
        float4 v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = ldcg_v4(B.u + (size_t)col[sg * 16 + e] * B.stride + 4 * q);
#pragma unroll
        for (int e = 0; e < 16; ++e) { a.x += v[e].x; a.y += v[e].y; a.z += v[e].z; a.w += v[e].w; }
      }
      tot.x += a.x; tot.y += a.y; tot.z += a.z; tot.w += a.w;
    }
    out[tid] = tot;
    __syncthreads();
  }
  return out[(tid + 1) % THREADS].x;
}

// ---- T12: latency of dependent fp64 / fp32 operations in one warp (what the scalar phase of the CG kernel is made of) ----
__device__ float t_chain(const Bufs& B, int kind) {
  if (threadIdx.x >= 32 || blockIdx.x != 0) return 0.f;
  double d = 1.0 + threadIdx.x, e = 1e-9;
  float f = 1.0f + threadIdx.x, g = 1e-6f;
  const long long t0 = clock64();
  for (int k = 0; k < B.reps; ++k) {
    if (kind == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) d = d + e;            // DADD chain
    } else if (kind == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) d = fma(d, 1.0000001, e);  // DFMA chain
    } else if (kind == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) d = fmax(d, e * (double)i) + e;  // DSETP/DMNMX + DADD
    } else if (kind == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f = fmaf(f, 1.0000001f, g);  // FFMA chain
    } else if (kind == 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) d = d + __shfl_xor_sync(0xffffffffu, d, 1 << (i & 3));  // double butterfly step
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) f = f + __shfl_xor_sync(0xffffffffu, f, 1 << (i & 3));  // float butterfly step
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) B.sink[1] = (float)(t1 - t0) / (16.f * B.reps);
  return (float)d + f;
}

__global__ void __launch_bounds__(THREADS, 1) bench_kernel(Bufs B) {
  float acc = 0.f;
  switch (B.mode) {
    case 1: t_pingpong(B); break;
    case 2: t_flags_mailbox(B); break;
    case 3: t_flags_shared(B); break;
    case 4: t_counter(B); break;
    case 5: acc = t_reduce_twohop(B); break;
    case 6: acc = t_reduce_onehop(B); break;
    case 7: t_flags_cluster(B); break;
    case 8: acc = t_reduce_cluster(B); break;
    case 9: t_gridsync(B); break;
    case 10: t_clusterbar(B); break;
    case 11: acc = t_gather(B); break;
    case 18: acc = t_reduce_barrier_read(B); break;
    case 19: acc = t_reduce_onehop_cols(B); break;
    case 20: acc = t_reduce_barrier_cols(B); break;
    case 12: case 13: case 14: case 15: case 16: case 17: acc = t_chain(B, B.mode - 12); break;
  }
  if (acc == -1.f) B.sink[0] = acc;
}

static double run(Bufs B, int grid, int cluster, size_t bytes_to_clear, void* clear_base) {
  CK(cudaMemset(clear_base, 0, bytes_to_clear));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = 0;
  cudaLaunchAttribute at[2];
  int na = 0;
  at[na].id = cudaLaunchAttributeCooperative;
  at[na].val.cooperative = 1;
  ++na;
  if (cluster > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = cluster;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  cudaError_t rc = cudaLaunchKernelEx(&cfg, bench_kernel, B);
  if (rc != cudaSuccess) {
    cudaGetLastError();
    return -1.0;
  }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return 1e6 * ms / B.reps;  // ns per exchange
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  printf("# %s, %d SMs, reps per launch 2000, %d threads per CTA\n", prop.name, sms, THREADS);
  char* base;
  const size_t bytes = sizeof(unsigned) * MAXG * MAXG + sizeof(unsigned long long) * NCOL * MAXG * 4 + 1024;
  CK(cudaMalloc(&base, bytes));
  Bufs B;
  B.flags = (unsigned*)base;
  B.words = (unsigned long long*)(base + sizeof(unsigned) * MAXG * MAXG);
  B.results = B.words + NCOL * MAXG;
  B.counter = (unsigned*)(B.words + NCOL * MAXG * 4);
  B.sink = (float*)(B.counter + 64);
  B.reps = 2000;
  CK(cudaMalloc(&B.u, sizeof(float) * GM * 16));
  CK(cudaMemset(B.u, 0, sizeof(float) * GM * 16));
  B.stride = 16;
  B.seg = 8;
  CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  // how many clusters of each size can be co-resident
  for (int cs : {2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64);
    cfg.blockDim = dim3(THREADS);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int ncl = 0;
    cudaError_t rc = cudaOccupancyMaxActiveClusters(&ncl, bench_kernel, &cfg);
    printf("# max co-resident clusters of %2d CTAs (1024 threads, no dynamic smem): %d (%s)\n", cs, ncl, cudaGetErrorString(rc));
    if (rc != cudaSuccess) cudaGetLastError();
  }
  struct T {
    const char* name;
    int mode, grid, cluster, fence, sleep_ns;
  };
  const int g8 = (sms / 8) * 8, g4 = (sms / 4) * 4;
  T tests[] = {
      {"pingpong 2 CTAs (round trip)", 1, sms, 1, 0, 0},
      {"pingpong 2 CTAs, nanosleep 20", 1, sms, 1, 0, 20},
      {"flags mailbox all-to-all, no fence", 2, sms, 1, 0, 0},
      {"flags mailbox all-to-all, 2 fences (= CG E1)", 2, sms, 1, 1, 0},
      {"flags mailbox, 2 fences, nanosleep 32", 2, sms, 1, 1, 32},
      {"flags mailbox, 2 fences, G=128", 2, 128, 1, 1, 0},
      {"flags mailbox, 2 fences, G=64", 2, 64, 1, 1, 0},
      {"flags mailbox, 2 fences, G=16", 2, 16, 1, 1, 0},
      {"flags shared lines, no fence", 3, sms, 1, 0, 0},
      {"flags shared lines, 2 fences", 3, sms, 1, 1, 0},
      {"flags shared lines, 2 fences, nanosleep 32", 3, sms, 1, 1, 32},
      {"atomic counter barrier (red.release + ld.acquire)", 4, sms, 1, 0, 0},
      {"atomic counter barrier, nanosleep 32", 4, sms, 1, 0, 32},
      {"cg grid.sync()", 9, sms, 1, 0, 0},
      {"reduce two-hop owners (= CG E2)", 5, sms, 1, 0, 0},
      {"reduce two-hop owners, nanosleep 32", 5, sms, 1, 0, 32},
      {"reduce two-hop owners, G=64", 5, 64, 1, 0, 0},
      {"reduce: counter barrier + coalesced read + barrier", 18, sms, 1, 0, 0},
      {"reduce: barrier + warp-per-column read + barrier (= CG now)", 20, sms, 1, 0, 0},
      {"reduce one-hop, tagged doubles, warp per column", 19, sms, 1, 0, 0},
      {"reduce one-hop all-read", 6, sms, 1, 0, 0},
      {"reduce one-hop all-read, nanosleep 32", 6, sms, 1, 0, 32},
      {"reduce one-hop all-read, G=64", 6, 64, 1, 0, 0},
      {"cluster barrier alone, cluster 8", 10, g8, 8, 0, 0},
      {"cluster barrier alone, cluster 4", 10, g4, 4, 0, 0},
      {"cluster barrier alone, cluster 2", 10, sms, 2, 0, 0},
      {"flags cluster-hier, cluster 8, no fence", 7, g8, 8, 0, 0},
      {"flags cluster-hier, cluster 8, fences", 7, g8, 8, 1, 0},
      {"flags cluster-hier, cluster 4, fences", 7, g4, 4, 1, 0},
      {"flags cluster-hier, cluster 2, fences", 7, sms, 2, 1, 0},
      {"reduce cluster-hier, cluster 8", 8, g8, 8, 0, 0},
      {"reduce cluster-hier, cluster 4", 8, g4, 4, 0, 0},
      {"reduce cluster-hier, cluster 2", 8, sms, 2, 0, 0},
      {"reduce cluster-hier, cluster 16 (non-portable)", 8, (sms / 16) * 16, 16, 0, 0},
  };
  {
    const char* names[] = {"DADD", "DFMA", "DMNMX+DADD", "FFMA", "double shfl_xor + DADD", "float shfl_xor + FADD"};
    for (int k = 0; k < 6; ++k) {
      B.mode = 12 + k;
      run(B, 1, 1, bytes, base);
      float cyc = 0.f;
      CK(cudaMemcpy(&cyc, B.sink + 1, sizeof(float), cudaMemcpyDeviceToHost));
      printf("dependent chain, one warp: %-24s %7.1f cycles per step\n", names[k], cyc);
    }
  }
  for (int stride : {12, 16})
    for (int seg : {4, 8, 16}) if (argc <= 1)
      for (int grid : {sms, 1}) {
        B.mode = 11;
        B.stride = stride;
        B.seg = seg;
        const double ns = run(B, grid, 1, bytes, base);
        printf("gather 2560 rows x 48 B per CTA, row stride %2d floats, %2d loads per thread, grid %3d : %9.1f ns per pass\n", stride, seg,
               grid, ns);
      }
  const bool quick = argc > 1;
  for (const T& t : tests) {
    if (quick && t.mode != 6 && t.mode != 18 && t.mode != 4 && t.mode != 19 && t.mode != 20) continue;
    B.mode = t.mode;
    B.fence = t.fence;
    B.sleep_ns = t.sleep_ns;
    int grid = t.grid;
    double ns = -1.0;
    // cooperative + cluster launches need every cluster co-resident: shrink the grid until the launch is accepted
    for (int tries = 0; tries < 12 && ns < 0; ++tries) {
      ns = run(B, grid, t.cluster, bytes, base);
      if (ns < 0) grid -= t.cluster > 1 ? t.cluster : 1;
    }
    printf("%-52s grid %3d cluster %2d : %9.1f ns per exchange\n", t.name, grid, t.cluster, ns);
    fflush(stdout);
  }
  return 0;
}

// K4, on-chip variant -- multi-right-hand-side Jacobi-preconditioned CG whose state never leaves the SM.
// (replaces spsolve at GLL.py:53 / GLL.py:93; stopping rule and per-column freeze of stable_conjgrad, GLL.py:247-276)
//
// Every CTA owns a block of rows for the whole solve and keeps x, r, p, s = Ap, w = Au (u = r/diag), 1/diag and -- when
// it fits -- its slice of the CSR in SHARED MEMORY across iterations.  The only global traffic per iteration is
//   * the preconditioned residual u, published by its owner and gathered by the CTAs whose rows reference it (L2), and
//   * 3*lp dot-product partials per CTA.
// Recurrences are Chronopoulos-Gear's single-reduction CG (one SpMV, one reduction of <r,u>, <w,u>, <r,r> per iteration):
//     p = u + b p ;  s = w + b s ;  x += a p ;  r -= a s ;  u = r/diag ;  w = A u
//     b = g'/g ,  a = g' / (d - b g'/a_old)      with g = <r,u>, d = <w,u>
// so an iteration has two grid-wide exchanges instead of textbook CG's three:
//   E1  "u is published": one release-store of an epoch flag per CTA, every CTA acquires all flags (u is double buffered)
//   E2  dot products: each CTA stores {fp32 partial, epoch} as ONE 64-bit word per column, so the value arrives with
//       its flag in a single L2 round trip and no fence is needed; the column's owner CTA sums the G words and
//       publishes {sum, epoch}, which is all the other CTAs poll (fan-in G per column instead of G*G).
// Sums over CTAs are taken in double in a fixed order and every CTA reads the same published value, so all CTAs take
// identical branches and the result
// is bit-reproducible (no floating-point atomics).  With a single CTA (small systems: the 512-row minibatch solves) the
// exchanges degenerate to __syncthreads and nothing but the CSR and the answer touches global memory.
#include <math.h>

#include "cg_common.cuh"

namespace gll {
namespace {

constexpr int CR_THREADS = 1024;
constexpr int CR_WARPS = CR_THREADS / 32;
constexpr int CR_ROWS_ILP = 4;          // rows a warp gathers for at once (independent L2 loads in flight)
constexpr size_t CR_SMEM_BUDGET = 200 * 1024;

struct CrParams {
  CgParams cg;
  float* ubuf;                // [2][m*lp] published u
  unsigned* flags;            // [grid][grid] E1 mailboxes: flags[d][b] = epoch of CTA b's latest published u
  unsigned long long* words;  // [3*lp][grid] E2 {partial, epoch}
  unsigned long long* results;  // [grid][3*l] E2 mailboxes {sum over CTAs, epoch}, written by the column's owner CTA
  int rows_cap;               // rows per CTA
  int csr_cap;                // nnz a CTA can cache in shared memory
  unsigned long long* trace;  // debug: [grid][16 passes][8 phases] %globaltimer stamps (NULL = off)
};

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// lane = q*S + s with S a power of two: xor butterfly inside each group of S lanes (every lane of the group gets the sum)
__device__ __forceinline__ float4 reduce_slots(float4 a, int S) {
#pragma unroll 1
  for (int o = S >> 1; o >= 1; o >>= 1) {
    a.x += __shfl_xor_sync(FULL, a.x, o);
    a.y += __shfl_xor_sync(FULL, a.y, o);
    a.z += __shfl_xor_sync(FULL, a.z, o);
    a.w += __shfl_xor_sync(FULL, a.w, o);
  }
  return a;
}

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x);
  acc.y = fmaf(w, v.y, acc.y);
  acc.z = fmaf(w, v.z, acc.z);
  acc.w = fmaf(w, v.w, acc.w);
}
__device__ __forceinline__ void dot4(float4& acc, const float4& a, const float4& b) {
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.y = fmaf(a.y, b.y, acc.y);
  acc.z = fmaf(a.z, b.z, acc.z);
  acc.w = fmaf(a.w, b.w, acc.w);
}
__device__ __forceinline__ float4 scale4(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 warp_sum4(float4 a) {
  a.x = warp_sum(a.x);
  a.y = warp_sum(a.y);
  a.z = warp_sum(a.z);
  a.w = warp_sum(a.w);
  return a;
}

__device__ __forceinline__ void cr_stamp(const CrParams& R, int pass, int phase, int who = 0) {
  if (R.trace != nullptr && threadIdx.x == who && pass < 16) {
    const unsigned long long t = (unsigned long long)clock64();  // SM cycle counter: cheap to read
    R.trace[((size_t)blockIdx.x * 16 + pass) * 8 + phase] = t;
  }
}

constexpr int CR_POLL = 5;  // 32 * CR_POLL >= max grid (148 SMs)

struct Ctrl {  // written by warp 0, read by everybody after a __syncthreads
  int stop;
  int pad[3];
};

template <bool SINGLE>
__global__ void __launch_bounds__(CR_THREADS, 1) cg_resident_kernel(CrParams R) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const CgParams& P = R.cg;
  const int lp = P.lp, Q = lp >> 2;
  int S = 1;  // neighbour slots per class quad: largest power of two with Q*S <= 32
  while (2 * S * Q <= 32) S <<= 1;
  const int G = gridDim.x, b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q_idx = lane / S, s_idx = lane - q_idx * S;
  const bool lane_on = lane < S * Q;
  const int q_ld = lane_on ? q_idx : 0;  // inactive lanes still issue (harmless) loads
  const int NCH = max(1, CR_WARPS / Q);  // row chunks of the dot-product phase (one warp per (chunk, class quad))

  const int row0 = min(P.m, b * R.rows_cap);
  const int rows = min(P.m, row0 + R.rows_cap) - row0;
  const size_t vec = (size_t)R.rows_cap * lp;

  // ---- shared memory carve-up (sizes mirrored by fixed_smem below) ----
  float* xs = reinterpret_cast<float*>(sm_raw);
  float* rs = xs + vec;
  float* ps = rs + vec;
  float* ss = ps + vec;
  float* ws = ss + vec;
  float* dg = ws + vec;               // diag          [rows_cap]
  float* dinv = dg + R.rows_cap;      // 1/diag        [rows_cap]
  int* lptr = reinterpret_cast<int*>(dinv + R.rows_cap);  // local CSR pointers [rows_cap + 1]
  double* red = reinterpret_cast<double*>(sm_raw + align_up((size_t)((char*)(lptr + R.rows_cap + 1) - (char*)sm_raw), 16));  // [3*lp] g', d, rr
  float* inv_g_old = reinterpret_cast<float*>(red + 3 * lp);  // [lp]  1/<r,u> of the previous pass
  float* inv_a_old = inv_g_old + lp;                           // [lp]  1/alpha of the previous pass
  float* wpart = inv_a_old + lp;                               // [3][CR_WARPS][lp]  per-chunk partial dots
  float* alpha = wpart + 3 * CR_WARPS * lp;              // [lp]
  float* beta = alpha + lp;                              // [lp]
  int* frozen = reinterpret_cast<int*>(beta + lp);       // [lp]
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(frozen + lp);     // 16 bytes
  int* ccol = reinterpret_cast<int*>(ctrl) + 4;          // [csr_cap]
  float* cval = reinterpret_cast<float*>(ccol + R.csr_cap);

  const int nnz0 = (rows > 0) ? __ldg(P.ptr + row0) : 0;
  const int nnz_mine = (rows > 0) ? __ldg(P.ptr + row0 + rows) - nnz0 : 0;
  const bool cached = nnz_mine <= R.csr_cap;
  const int* col = cached ? ccol : P.col + nnz0;
  const float* val = cached ? cval : P.val + nnz0;

#pragma unroll 1
  for (int i = tid; i <= rows; i += CR_THREADS) lptr[i] = __ldg(P.ptr + row0 + i) - nnz0;
#pragma unroll 1
  for (int i = tid; i < rows; i += CR_THREADS) {
    const float d = __ldg(P.diag + row0 + i);
    dg[i] = d;
    dinv[i] = 1.f / d;
  }
  if (cached)
#pragma unroll 1
    for (int e = tid; e < nnz_mine; e += CR_THREADS) {
      ccol[e] = __ldg(P.col + nnz0 + e);
      cval[e] = __ldg(P.val + nnz0 + e);
    }
#pragma unroll 1
  for (int c = tid; c < lp; c += CR_THREADS) {
    inv_g_old[c] = 1.f;
    inv_a_old[c] = 1.f;
    frozen[c] = 0;
  }
  __syncthreads();
  // x = 0, r = b, p = s = 0; publish u0 = r/diag
  const int items = rows * Q;
#pragma unroll 1
  for (int it = tid; it < items; it += CR_THREADS) {
    const int i = it / Q, q = it - i * Q;
    const size_t o = (size_t)i * lp + 4 * q;
    const float4 bq = __ldg(reinterpret_cast<const float4*>(P.rhs + (size_t)(row0 + i) * lp + 4 * q));
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(xs + o) = z;
    *reinterpret_cast<float4*>(rs + o) = bq;
    *reinterpret_cast<float4*>(ps + o) = z;
    *reinterpret_cast<float4*>(ss + o) = z;
    if (!SINGLE) *reinterpret_cast<float4*>(R.ubuf + (size_t)(row0 + i) * lp + 4 * q) = scale4(bq, dinv[i]);
  }

  double tol2 = 0.0;  // squared absolute tolerance; meaningful in warp 0 only
  int iter = 0;

  while (true) {
    // ================= E1: u of this pass is visible everywhere =================
    const unsigned epoch = (unsigned)iter + 1u;
    __syncthreads();
    cr_stamp(R, iter, 0);
    if (!SINGLE) {
      // Mailboxes: CTA b tells every CTA d "my u is published" by writing flags[d][b]; each CTA polls only its own row
      // of the table, so no line is spun on by more than one SM (148 SMs spinning on shared lines delay the very
      // stores they wait for by microseconds).
      if (warp == 0) {
        __threadfence();  // the CTA's u stores (ordered before this warp by the barrier) become visible first
#pragma unroll
        for (int k = 0; k < CR_POLL; ++k)
          if (lane + 32 * k < G) st_relaxed_u32(R.flags + (size_t)(lane + 32 * k) * G + b, epoch);
      }
      if (warp == 1) {  // all sources polled concurrently with relaxed loads; ONE acquire fence at the end
        const unsigned* mine = R.flags + (size_t)b * G;
        bool done[CR_POLL];
#pragma unroll
        for (int k = 0; k < CR_POLL; ++k) done[k] = lane + 32 * k >= G;
        bool all;
        do {
          all = true;
#pragma unroll
          for (int k = 0; k < CR_POLL; ++k)
            if (!done[k]) {
              done[k] = ld_relaxed_u32(mine + lane + 32 * k) >= epoch;
              all &= done[k];
            }
        } while (!all);
        __threadfence();
      }
      __syncthreads();
    }
    cr_stamp(R, iter, 1);
    // ================= w = A u for my rows =================
    if (SINGLE) {
      // every referenced row is local: gather straight from shared memory, one thread per (row, class quad)
#pragma unroll 1
      for (int it = tid; it < items; it += CR_THREADS) {
        const int i = it / Q, q = it - i * Q;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int e = lptr[i]; e < lptr[i + 1]; ++e) {
          const int j = col[e];
          fma4(a, val[e] * dinv[j], *reinterpret_cast<const float4*>(rs + (size_t)j * lp + 4 * q));
        }
        const size_t o = (size_t)i * lp + 4 * q;
        const float4 u4 = scale4(*reinterpret_cast<const float4*>(rs + o), dinv[i]);
        const float dgi = dg[i];
        *reinterpret_cast<float4*>(ws + o) = make_float4(fmaf(dgi, u4.x, -a.x), fmaf(dgi, u4.y, -a.y), fmaf(dgi, u4.z, -a.z), fmaf(dgi, u4.w, -a.w));
      }
    } else {
      // warp per row, lane = (neighbour slot, class quad); CR_ROWS_ILP rows x 2 slots per lane = 8 independent
      // 128-bit L2 gathers in flight before the first use
      const float* ub = R.ubuf + (size_t)(iter & 1) * P.m * lp;
#pragma unroll 1
      for (int ibase = warp; ibase < rows; ibase += CR_WARPS * CR_ROWS_ILP) {
        float4 a[CR_ROWS_ILP];
        int e[CR_ROWS_ILP], e_end[CR_ROWS_ILP];
        bool more = false;
#pragma unroll
        for (int t = 0; t < CR_ROWS_ILP; ++t) {
          const int i = ibase + t * CR_WARPS;
          a[t] = make_float4(0.f, 0.f, 0.f, 0.f);
          const bool ok = lane_on && i < rows;
          e[t] = ok ? lptr[i] + s_idx : 0;
          e_end[t] = ok ? lptr[i + 1] : 0;
          more |= e[t] < e_end[t];
        }
#pragma unroll 1
        while (more) {
          float4 u0[CR_ROWS_ILP], u1[CR_ROWS_ILP];
          float w0[CR_ROWS_ILP], w1[CR_ROWS_ILP];
#pragma unroll
          for (int t = 0; t < CR_ROWS_ILP; ++t) {
            const bool v0 = e[t] < e_end[t], v1 = e[t] + S < e_end[t];
            w0[t] = v0 ? val[e[t]] : 0.f;
            w1[t] = v1 ? val[e[t] + S] : 0.f;
            const int j0 = v0 ? col[e[t]] : row0, j1 = v1 ? col[e[t] + S] : row0;  // row0: a harmless valid address
            u0[t] = ldcg4(ub + (size_t)j0 * lp + 4 * q_ld);
            u1[t] = ldcg4(ub + (size_t)j1 * lp + 4 * q_ld);
          }
          more = false;
#pragma unroll
          for (int t = 0; t < CR_ROWS_ILP; ++t) {
            fma4(a[t], w0[t], u0[t]);
            fma4(a[t], w1[t], u1[t]);
            e[t] += 2 * S;
            more |= e[t] < e_end[t];
          }
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < CR_ROWS_ILP; ++t) {
          const int i = ibase + t * CR_WARPS;
          const float4 at = reduce_slots(a[t], S);
          if (i < rows && lane_on && s_idx == 0) {
            const size_t o = (size_t)i * lp + 4 * q_idx;
            const float4 u4 = scale4(*reinterpret_cast<const float4*>(rs + o), dinv[i]);
            const float dgi = dg[i];
            *reinterpret_cast<float4*>(ws + o) = make_float4(fmaf(dgi, u4.x, -at.x), fmaf(dgi, u4.y, -at.y), fmaf(dgi, u4.z, -at.z), fmaf(dgi, u4.w, -at.w));
          }
        }
      }
    }
    __syncthreads();
    cr_stamp(R, iter, 2);
    // ================= partial <r,u>, <w,u>, <r,r>: one warp per (row chunk, class quad), fixed summation order ======
#pragma unroll 1
    for (int job = warp; job < NCH * Q; job += CR_WARPS) {
      const int ch = job / Q, q = job - ch * Q;
      const int i_lo = (int)((long long)rows * ch / NCH), i_hi = (int)((long long)rows * (ch + 1) / NCH);
      float4 d_ru = make_float4(0.f, 0.f, 0.f, 0.f), d_wu = d_ru, d_rr = d_ru;
#pragma unroll 1
      for (int i = i_lo + lane; i < i_hi; i += 32) {
        const size_t o = (size_t)i * lp + 4 * q;
        const float4 r4 = *reinterpret_cast<const float4*>(rs + o), w4 = *reinterpret_cast<const float4*>(ws + o);
        const float4 u4 = scale4(r4, dinv[i]);
        dot4(d_ru, r4, u4);
        dot4(d_wu, w4, u4);
        dot4(d_rr, r4, r4);
      }
      d_ru = warp_sum4(d_ru);
      d_wu = warp_sum4(d_wu);
      d_rr = warp_sum4(d_rr);
      if (lane == 0) {
        *reinterpret_cast<float4*>(wpart + ((size_t)0 * CR_WARPS + ch) * lp + 4 * q) = d_ru;
        *reinterpret_cast<float4*>(wpart + ((size_t)1 * CR_WARPS + ch) * lp + 4 * q) = d_wu;
        *reinterpret_cast<float4*>(wpart + ((size_t)2 * CR_WARPS + ch) * lp + 4 * q) = d_rr;
      }
    }
    __syncthreads();
    cr_stamp(R, iter, 3);
    // ================= E2: dot products over all CTAs =================
    if (tid < lp) {  // thread cc sums the chunk partials of its class column for the three dot products
      const int cc = tid;
#pragma unroll 1
      for (int v = 0; v < 3; ++v) {
        double t = 0.0;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) t += (double)wpart[((size_t)v * CR_WARPS + ch) * lp + cc];
        if (SINGLE || cc >= P.l) {
          red[v * lp + cc] = (cc < P.l) ? t : 0.0;
        } else {
          const unsigned long long word = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint((float)t);
          st_relaxed_u64(R.words + (size_t)(v * lp + cc) * G + b, word);
        }
      }
    }
    if (!SINGLE) {
      // Padded class columns are identically zero: only the 3*l real columns travel.  Column cr is summed by ONE owner
      // CTA (cr mod G), which publishes {sum, epoch}; everybody else polls just those 3*l result words.  (All CTAs
      // polling all G partials of every column is an L2 hot spot: G*G*3l reads of a few hundred lines per round.)
      const int ncols = 3 * P.l;
#pragma unroll 1
      for (int k = warp; b + k * G < ncols; k += CR_WARPS) {
        const int cr = b + k * G;
        const int v_of = (cr >= 2 * P.l) ? 2 : (cr >= P.l ? 1 : 0);
        const int c = v_of * lp + (cr - v_of * P.l);
        const unsigned long long* wbase = R.words + (size_t)c * G;
        unsigned long long word[CR_POLL];
        bool done[CR_POLL];
#pragma unroll
        for (int kk = 0; kk < CR_POLL; ++kk) {
          done[kk] = lane + 32 * kk >= G;
          word[kk] = 0ull;
        }
        bool all;
        do {  // all CTAs' words polled concurrently; the value arrives with its epoch tag
          all = true;
#pragma unroll
          for (int kk = 0; kk < CR_POLL; ++kk)
            if (!done[kk]) {
              word[kk] = ld_relaxed_u64(wbase + lane + 32 * kk);
              done[kk] = (unsigned)(word[kk] >> 32) == epoch;
              all &= done[kk];
            }
        } while (!all);
        double t = 0.0;
#pragma unroll
        for (int kk = 0; kk < CR_POLL; ++kk)
          if (lane + 32 * kk < G) t += (double)__uint_as_float((unsigned)word[kk]);
        t = warp_sum(t);  // xor butterfly: fixed order
        const unsigned long long out = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint((float)t);
#pragma unroll
        for (int kk = 0; kk < CR_POLL; ++kk)  // one private copy per CTA: results[d][cr]
          if (lane + 32 * kk < G) st_relaxed_u64(R.results + (size_t)(lane + 32 * kk) * ncols + cr, out);
      }
      if (warp == CR_WARPS - 1) {
        const unsigned long long* mine = R.results + (size_t)b * ncols;
#pragma unroll 1
        for (int cr = lane; cr < ncols; cr += 32) {
          unsigned long long word;
          do {
            word = ld_relaxed_u64(mine + cr);
          } while ((unsigned)(word >> 32) != epoch);
          const int v_of = (cr >= 2 * P.l) ? 2 : (cr >= P.l ? 1 : 0);
          red[v_of * lp + (cr - v_of * P.l)] = (double)__uint_as_float((unsigned)word);
        }
        __syncwarp();
        cr_stamp(R, iter, 4, (CR_WARPS - 1) * 32);
      }
    }
    __syncthreads();

    // ================= scalars (warp 0; every CTA computes identical values from identical inputs) =================
    if (warp == 0) {
      if (iter == 0) {  // P.tol < 0: relative to the largest right-hand-side column norm (r = b at this point)
        double mx = 0.0;
#pragma unroll 1
        for (int c = lane; c < lp; c += 32) mx = fmax(mx, red[2 * lp + c]);
#pragma unroll 1
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
        tol2 = (P.tol < 0.f) ? (double)P.tol * (double)P.tol * mx : (double)P.tol * (double)P.tol;
      }
      double mx_all = 0.0, mx_live = 0.0;
      int bad = 0;
#pragma unroll 1
      for (int c = lane; c < lp; c += 32) {
        const double v = red[2 * lp + c];
        bad |= (!(v == v) || v > 1.0e300) ? 1 : 0;
        mx_all = fmax(mx_all, v);
        if (!frozen[c]) mx_live = fmax(mx_live, v);  // columns that broke down are frozen for good
      }
#pragma unroll 1
      for (int o = 16; o > 0; o >>= 1) {
        mx_all = fmax(mx_all, __shfl_xor_sync(FULL, mx_all, o));
        mx_live = fmax(mx_live, __shfl_xor_sync(FULL, mx_live, o));
        bad |= __shfl_xor_sync(FULL, bad, o);
      }
      const bool stop = bad || mx_live <= tol2 || iter >= P.max_iter;
      cr_stamp(R, iter, 6);
      if (lane == 0) {
        ctrl->stop = stop ? 1 : 0;
        if (stop && b == 0) {
          if (P.iters_out) *P.iters_out = iter;
          if (P.resid_out) *P.resid_out = sqrtf((float)mx_all);
          if (P.status_out) {
            int st = 0;
            if (bad) st |= GLL_STATUS_NONFINITE;
            if (!bad && !(mx_all <= tol2)) st |= GLL_STATUS_CG_NOT_CONVERGED;
            if (st) atomicOr(P.status_out, st);
          }
        }
      }
      if (!stop) {
#pragma unroll 1
        for (int c = lane; c < lp; c += 32) {
          // quotients in fp32 (alpha and beta are fp32 anyway); the cancellation-prone difference in fp64
          const double g_new = red[c], d_new = red[lp + c], rr = red[2 * lp + c];
          float al = 0.f, be = 0.f;
          if (!frozen[c] && rr > tol2) {
            const float bb = (iter == 0) ? 0.f : (float)g_new * inv_g_old[c];
            const double den = d_new - (double)bb * g_new * (double)inv_a_old[c];
            if (den > 0.0 && g_new > 0.0) {
              al = (float)g_new / (float)den;
              be = bb;
              inv_a_old[c] = 1.f / al;
              inv_g_old[c] = 1.f / (float)g_new;
            } else {
              frozen[c] = 1;  // breakdown (rounding at the fp32 floor): stop moving this column
            }
          }
          alpha[c] = al;
          beta[c] = be;
        }
      }
      cr_stamp(R, iter, 7);
    }
    __syncthreads();
    cr_stamp(R, iter, 5);
    if (ctrl->stop) break;
    ++iter;
    // ================= vector updates (all on chip), publish the new u =================
    float* unext = SINGLE ? nullptr : R.ubuf + (size_t)(iter & 1) * P.m * lp;
#pragma unroll 1
    for (int it = tid; it < items; it += CR_THREADS) {
      const int i = it / Q, q = it - i * Q;
      const size_t o = (size_t)i * lp + 4 * q;
      const float4 al = *reinterpret_cast<const float4*>(alpha + 4 * q);
      const float4 be = *reinterpret_cast<const float4*>(beta + 4 * q);
      const float di = dinv[i];
      float4 x4 = *reinterpret_cast<const float4*>(xs + o), r4 = *reinterpret_cast<const float4*>(rs + o);
      float4 p4 = *reinterpret_cast<const float4*>(ps + o), s4 = *reinterpret_cast<const float4*>(ss + o);
      const float4 w4 = *reinterpret_cast<const float4*>(ws + o);
      p4.x = fmaf(be.x, p4.x, r4.x * di); p4.y = fmaf(be.y, p4.y, r4.y * di); p4.z = fmaf(be.z, p4.z, r4.z * di); p4.w = fmaf(be.w, p4.w, r4.w * di);
      s4.x = fmaf(be.x, s4.x, w4.x); s4.y = fmaf(be.y, s4.y, w4.y); s4.z = fmaf(be.z, s4.z, w4.z); s4.w = fmaf(be.w, s4.w, w4.w);
      x4.x = fmaf(al.x, p4.x, x4.x); x4.y = fmaf(al.y, p4.y, x4.y); x4.z = fmaf(al.z, p4.z, x4.z); x4.w = fmaf(al.w, p4.w, x4.w);
      r4.x = fmaf(-al.x, s4.x, r4.x); r4.y = fmaf(-al.y, s4.y, r4.y); r4.z = fmaf(-al.z, s4.z, r4.z); r4.w = fmaf(-al.w, s4.w, r4.w);
      *reinterpret_cast<float4*>(xs + o) = x4;
      *reinterpret_cast<float4*>(rs + o) = r4;
      *reinterpret_cast<float4*>(ps + o) = p4;
      *reinterpret_cast<float4*>(ss + o) = s4;
      if (!SINGLE) *reinterpret_cast<float4*>(unext + (size_t)(row0 + i) * lp + 4 * q) = scale4(r4, di);
    }
  }

  // ---- write the answer ----
#pragma unroll 1
  for (int it = tid; it < items; it += CR_THREADS) {
    const int i = it / Q, q = it - i * Q;
    *reinterpret_cast<float4*>(P.x + (size_t)(row0 + i) * lp + 4 * q) = *reinterpret_cast<const float4*>(xs + (size_t)i * lp + 4 * q);
  }
}

size_t fixed_smem(int rows_cap, int lp) {
  size_t b = sizeof(float) * (5 * (size_t)rows_cap * lp + 2 * (size_t)rows_cap) + sizeof(int) * ((size_t)rows_cap + 1);
  b = align_up(b, 16);
  b += sizeof(double) * (3 * (size_t)lp) + sizeof(float) * (3 * (size_t)CR_WARPS * lp + 4 * lp) + sizeof(int) * lp + 16;
  return align_up(b, 16);
}

struct CrPlan {
  int ok, grid, rows_cap, csr_cap;
  size_t smem;
};

CrPlan plan(int m, int lp) {
  CrPlan p;
  p.ok = 0;
  const int sms = device_info().sms;
  int grid;
  if (fixed_smem(m, lp) + 8 * 1024 <= CR_SMEM_BUDGET * 3 / 4)
    grid = 1;  // whole system in one SM: no grid-wide exchange at all
  else
    grid = max(2, min(sms, ceil_div(m, 96)));
  int rows_cap = ceil_div(m, grid);
  if (grid > 1 && fixed_smem(rows_cap, lp) + 4 * 1024 > CR_SMEM_BUDGET) {
    grid = sms;
    rows_cap = ceil_div(m, grid);
    if (fixed_smem(rows_cap, lp) + 4 * 1024 > CR_SMEM_BUDGET) return p;  // does not fit on chip: streaming kernel
  }
  p.grid = grid;
  p.rows_cap = rows_cap;
  const size_t fx = fixed_smem(rows_cap, lp);
  p.csr_cap = (int)((CR_SMEM_BUDGET - fx) / 8);
  p.smem = fx + (size_t)p.csr_cap * 8;
  p.ok = 1;
  return p;
}

}  // namespace

static unsigned long long* g_cg_trace = nullptr;
void cg_set_trace(void* buf) { g_cg_trace = (unsigned long long*)buf; }
void* cg_get_trace() { return g_cg_trace; }

size_t cg_resident_ws_bytes(int m, int lp) {
  const size_t sms = (size_t)device_info().sms;
  return align_up(sizeof(float) * 2 * (size_t)m * lp, 256) + align_up(sizeof(unsigned) * sms * sms, 256) +
         align_up(sizeof(unsigned long long) * 3 * (size_t)lp * sms * 2, 256) + 256;
}

int cg_resident_try(const CgParams& P, void* scratch, cudaStream_t st) {
  const CrPlan pl = plan(P.m, P.lp);
  if (!pl.ok) return 0;
  CrParams R;
  R.cg = P;
  Carver cv(scratch, cg_resident_ws_bytes(P.m, P.lp));
  R.ubuf = cv.take<float>(2 * (size_t)P.m * P.lp);
  const size_t sms = (size_t)device_info().sms;
  R.flags = cv.take<unsigned>(sms * sms);
  R.words = cv.take<unsigned long long>(3 * (size_t)P.lp * sms * 2);
  R.results = R.words + 3 * (size_t)P.lp * sms;
  R.rows_cap = pl.rows_cap;
  R.csr_cap = pl.csr_cap;
  R.trace = g_cg_trace;
  static bool attr_set = false;
  if (!attr_set) {
    GLL_CUDA_CHECK(cudaFuncSetAttribute(cg_resident_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CR_SMEM_BUDGET));
    GLL_CUDA_CHECK(cudaFuncSetAttribute(cg_resident_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CR_SMEM_BUDGET));
    attr_set = true;
  }
  void* args[] = {&R};
  if (pl.grid == 1) {
    GLL_PROF(KID_CG, st);
    cg_resident_kernel<true><<<1, CR_THREADS, pl.smem, st>>>(R);
    GLL_LAUNCH_CHECK();
  } else {
    // flags and words live side by side: one memset clears the epochs of both exchanges
    GLL_CUDA_CHECK(cudaMemsetAsync(R.flags, 0, (size_t)((char*)(R.results + 3 * (size_t)P.lp * sms) - (char*)R.flags), st));
    GLL_PROF(KID_CG, st);
    // cooperative launch: all CTAs are co-resident, which the flag exchanges rely on
    GLL_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)cg_resident_kernel<false>, dim3(pl.grid), dim3(CR_THREADS), args, pl.smem, st));
  }
  return 1;
}

}  // namespace gll

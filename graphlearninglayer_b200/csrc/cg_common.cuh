// Shared by the two CG kernels: cg.cu (streaming: vectors in global memory, any size) and cg_resident.cu
// (vectors and the CSR slice of every CTA stay in shared memory across iterations).
#pragma once
#include "common.cuh"

namespace gll {

constexpr int CG_THREADS = 1024;
constexpr int CG_WARPS = CG_THREADS / 32;
constexpr int CG_MAX_LP = 128;

struct CgParams {
  const int* ptr;
  const int* col;
  const float* val;
  const float* diag;
  const float* rhs;
  float* x;
  float* r;
  float* p;
  float* ap;
  double* partial;    // [2 buffers][2*lp columns][grid]
  unsigned* barrier;  // zeroed before launch
  unsigned* ext_counter;  // optional [2], zero on entry: the on-chip kernel's grid barrier and exit ticket.  The kernel leaves
                          // both zero again, so one memset per forward (the info block) serves the forward AND the adjoint solve
  int m, l, lp, rows_per_block, max_iter;
  float tol;
  int* iters_out;
  float* resid_out;
  int* status_out;
  // Fused input / output conversions of the layer (on-chip kernels; cg_run runs the separate kernels for the others):
  const void* rhs_src;  // if set: the right-hand side is this [m][l] array (rhs_kind 1: fp32, 2: fp64) instead of rhs [m][lp]
  int rhs_kind;
  void* x_copy;         // if set: the solution is also written as an [m][l] array (fp64 if x_copy_f64, else fp32)
  int x_copy_f64;
  const void* rhs_scale;  // optional device scalar (fp64 if rhs_scale_f64, else fp32) multiplied into rhs_src: the upstream gradient
  int rhs_scale_f64;      // of a loss head whose d loss / d pred is rhs_src (losses.py: LaplaceLearningCELoss)
};

// right-hand side of class quad q of a row, from the padded fp32 array or straight from the caller's [m][l] array
__device__ __forceinline__ float4 cg_load_rhs4(const CgParams& P, int row, int q) {
  if (P.rhs_src == nullptr) return __ldg(reinterpret_cast<const float4*>(P.rhs + (size_t)row * P.lp + 4 * q));
  double sc = 1.0;
  if (P.rhs_scale != nullptr)
    sc = P.rhs_scale_f64 ? __ldg(reinterpret_cast<const double*>(P.rhs_scale)) : (double)__ldg(reinterpret_cast<const float*>(P.rhs_scale));
  float v[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int col = 4 * q + c;
    v[c] = 0.f;
    if (col < P.l)
      v[c] = (P.rhs_kind == 2) ? (float)(sc * __ldg(reinterpret_cast<const double*>(P.rhs_src) + (size_t)row * P.l + col))
                               : (float)(sc * (double)__ldg(reinterpret_cast<const float*>(P.rhs_src) + (size_t)row * P.l + col));
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void cg_store_copy4(const CgParams& P, int row, int q, const float4& x) {
  if (P.x_copy == nullptr) return;
  const float v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int col = 4 * q + c;
    if (col < P.l) {
      if (P.x_copy_f64)
        reinterpret_cast<double*>(P.x_copy)[(size_t)row * P.l + col] = (double)v[c];
      else
        reinterpret_cast<float*>(P.x_copy)[(size_t)row * P.l + col] = v[c];
    }
  }
}

// Resident variant (cg_resident.cu): returns 1 if it took the solve, 0 if the system does not fit on chip, <0 on error.
// scratch must hold cg_resident_ws_bytes(m, lp) bytes.
size_t cg_resident_ws_bytes(int m, int lp);
int cg_resident_try(const CgParams& P, void* scratch, cudaStream_t st);
void cg_set_trace(void* buf);
void* cg_get_trace();
// One-CTA kernel for minibatch-sized systems (cg_small.cu): 1 = took the solve, 0 = not small, < 0 = error.
int cg_small_try(const CgParams& P, cudaStream_t st);
// Cluster variant for the same minibatch-sized systems (cg_cluster.cu): one cluster of eight CTAs, exchanges through DSMEM.
int cg_cluster_try(const CgParams& P, cudaStream_t st);

}  // namespace gll

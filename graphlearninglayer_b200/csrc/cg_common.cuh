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
};

// Resident variant (cg_resident.cu): returns 1 if it took the solve, 0 if the system does not fit on chip, <0 on error.
// scratch must hold cg_resident_ws_bytes(m, lp) bytes.
size_t cg_resident_ws_bytes(int m, int lp);
int cg_resident_try(const CgParams& P, void* scratch, cudaStream_t st);
void cg_set_trace(void* buf);
void* cg_get_trace();
// One-CTA kernel for minibatch-sized systems (cg_small.cu): 1 = took the solve, 0 = not small, < 0 = error.
int cg_small_try(const CgParams& P, cudaStream_t st);

}  // namespace gll

// Host-side objects behind the opaque handles of include/hsolve_cuda.h.
#pragma once

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "hs_internal.h"
#include "hs_types.cuh"

#define CUDA_OK(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      throw hs_error(e__ == cudaErrorMemoryAllocation ? HS_ENOMEM : HS_ECUDA,                           \
                     std::string(#call) + ": " + cudaGetErrorString(e__));                              \
  } while (0)

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct hs_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t aux_stream = nullptr;  // look-ahead: the big trailing update overlaps the next block's panels
  cudaEvent_t ev_b = nullptr, ev_c2 = nullptr;
  int lookahead_max_fronts = 32;
  int max_cluster = 8;
  bool profile = false;
  long long launches = 0;
  void* gm_buf = nullptr;   // GMRES workspace (Krylov basis + work vectors), grow-only, reused across hs_gmres calls
  size_t gm_bytes = 0;
  int outer_block = 256;  // NB of the two-level blocked LU (HS_OUTER_BLOCK)
};

// ------------------------------------------------------------------------------------------------
// factorization object
// ------------------------------------------------------------------------------------------------
struct Level {
  int f0 = 0, f1 = 0;        // front range
  int max_n = 0, max_ni = 0, max_nb = 0;
  long long ioff0 = 0, ioff1 = 0;
  long long poff0 = 0, poff1 = 0;
  bool pseudo = false;
  std::vector<int> ni_sorted;  // ni of the fronts in this level (descending)
};

struct hs_fac {
  hs_ctx* ctx = nullptr;
  hs_dtype dtype = HS_F64;
  size_t esz = 8;
  int64_t n = 0, nnz = 0, nnodes = 0;
  hs_opts opts{};
  int64_t swlevel_resolved = 0, depth = 0;
  // host copies of the symbolic data (0-based)
  std::vector<int64_t> left, right, parent, level;
  std::vector<int64_t> iloc_ptr, iloc_idx, bloc_ptr, bloc_idx;
  std::vector<int> node_ni, node_nb, node2front;
  std::vector<Front> fronts;
  std::vector<Level> levels;  // deepest first, root last (+ pseudo front for a non-empty root boundary)
  int root_front = -1, pseudo_front = -1;
  // external leaves (subtree-per-GPU): front id → device buffer its Schur block is imported from
  std::vector<const void*> ext_src;
  std::vector<long long> ext_ld;
  long long pool_elems = 0, idx_total = 0, max_level_idx = 0;
  // device
  void* pool = nullptr;
  Front* d_fronts = nullptr;
  int *d_gidx = nullptr, *d_ipiv = nullptr, *d_rperm = nullptr, *d_cmap = nullptr, *d_own = nullptr, *d_pos = nullptr,
      *d_info = nullptr;
  long long *d_colptr = nullptr, *d_rowval = nullptr;
  void* d_nzval = nullptr;
  void *d_x = nullptr, *d_work = nullptr;
  // CSR image of A for the GMRES mat-vec (built on first use)
  long long *d_csr_ptr = nullptr, *d_csr_col = nullptr;
  void* d_csr_val = nullptr;
  int64_t rhs_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  hs_stats_t stats{};

  ~hs_fac() {
    cudaFree(pool); cudaFree(d_fronts); cudaFree(d_gidx); cudaFree(d_ipiv); cudaFree(d_rperm); cudaFree(d_cmap);
    cudaFree(d_own); cudaFree(d_pos); cudaFree(d_info); cudaFree(d_colptr); cudaFree(d_rowval); cudaFree(d_nzval);
    cudaFree(d_x); cudaFree(d_work); cudaFree(d_csr_ptr); cudaFree(d_csr_col); cudaFree(d_csr_val);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
};


// hs_panel_f64.cu / hs_panel_c64.cu
void hs_panel_setup_f64();
void hs_panel_setup_c64();
int hs_panel_width_f64(const hs_fac* f, int max_n, int nfronts);
int hs_panel_width_c64(const hs_fac* f, int max_n, int nfronts);
void hs_panel_launch_f64(hs_fac* f, int W, int f0, int nact, int j0, int m, cudaStream_t st);
void hs_panel_launch_c64(hs_fac* f, int W, int f0, int nact, int j0, int m, cudaStream_t st);
inline int hs_panel_width(const hs_fac* f, int max_n, int nfronts) {
  return f->dtype == HS_F64 ? hs_panel_width_f64(f, max_n, nfronts) : hs_panel_width_c64(f, max_n, nfronts);
}
inline void hs_panel_launch(hs_fac* f, int W, int f0, int nact, int j0, int m, cudaStream_t st) {
  if (f->dtype == HS_F64) hs_panel_launch_f64(f, W, f0, nact, j0, m, st); else hs_panel_launch_c64(f, W, f0, nact, j0, m, st);
}

// hs_solve.cu
void hs_solve_setup();
int hs_solve_block(hs_dtype dt);
void hs_solve_prep(hs_fac* f, const Level& L);
void hs_solve_run(hs_fac* f, int64_t nrhs, void* x, int which);  // which: 1 forward, 2 backward, 3 both

// hs_small.cu
int hs_small_max_n(hs_dtype dt);
void hs_small_factor(hs_fac* f, const Level& L);

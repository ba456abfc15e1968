/*
 * libhsolve_cuda — C ABI of the B200-native multifrontal nested-dissection factor + tree solve.
 *
 * This is the drop-in boundary for the hot path of bonevbs/HierarchicalSolvers.jl.  The reference has no
 * FFI of its own (it is plain Julia dispatch), so each entry point below names the Julia method whose
 * work it takes over; the Julia-side `ccall` binding a maintainer would add is in INTEGRATION.md and
 * hierarchicalsolvers.jl_b200/julia/HierarchicalSolversCUDA.jl.
 *
 * Conventions
 *   - every function returns an int32 status (HS_OK = 0); hs_last_error() gives the text of the last
 *     failure on the calling thread.  Nothing throws across this boundary.
 *   - all host inputs are copied during the call; the library owns every device allocation until the
 *     matching *_free / hs_destroy.
 *   - matrices are column-major; complex is interleaved (re, im) float64.
 *   - index arrays are int64 and carry an explicit `index_base` (Julia passes 1, C/Python 0).
 */
#ifndef HSOLVE_CUDA_H
#define HSOLVE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_VERSION 200

/* status codes — the Julia shim maps them to the exception the reference would raise */
enum {
  HS_OK = 0,
  HS_EARG = 1,      /* ArgumentError      (HierarchicalSolvers.jl:74-78, nesteddissection.jl:111)        */
  HS_EDIM = 2,      /* DimensionMismatch  (blockmatrix.jl:13-16,116-117, nesteddissection.jl:107)         */
  HS_ETREE = 3,     /* ErrorException     (factorization.jl:25: node with a single child)                */
  HS_ESINGULAR = 4, /* SingularException  (LAPACK getrf info > 0 behind every `\` in the reference)       */
  HS_ECUDA = 5,
  HS_ENOMEM = 6,
  HS_ENOTIMPL = 7,
  HS_ESIZE = 8      /* a front exceeds what the panel kernels cover                                      */
};

typedef enum { HS_F64 = 0, HS_C64 = 1 } hs_dtype;

/* `SolverOptions` (HierarchicalSolvers.jl:30-40), same field names, followed by this library's extensions.
 * `swlevel` is passed as the user gave it; negative values are resolved against the tree depth inside hs_factor as
 * factorization.jl:8 does.  A node with level ≤ swlevel and |bnd| ≥ swsize is compressed (factorization.jl:15):
 *   - its L and R become low-rank, truncated at atol/2, rtol/2 (:99-100, :171-209);
 *   - with `hss` != 0 (the reference's behaviour) its Schur complement S[perm,perm] is stored as an HSS matrix built by
 *     the adaptive randomized construction (`randcompress_adaptive`, :102-110) from the matrix-free operator of :228-249:
 *     cluster tree = bisection down to `leafsize` with the first split forced at |int_loc| (:109), `kest` + 10 Gaussian
 *     samples (kest < 0: ceil(rank(L)/2), :102-104), `stepsize` more whenever a detected rank saturates the sample
 *     count, truncation at atol, rtol.  Parents then assemble from the HSS-approximated blocks and take their low-rank
 *     Gauss transforms from the children's generators (:126-140, :184-209);
 *   - with `hss` == 0 S is evaluated exactly and kept dense (the limit of a zero HSS tolerance); `leafsize`, `kest`,
 *     `stepsize` then have no effect.
 * `c_tol` is validated and ignored, as in the reference (:97).
 * Accuracy floor: the low-rank Gauss transforms truncate through a pivoted Cholesky factorization of a Gram matrix, which
 * resolves singular values down to about 1e-7·‖block‖; atol/rtol below that (including 0) drive the ranks to
 * min(ni, nb) without a gain in accuracy.  The HSS construction uses Householder QR with column pivoting (no such floor). */
typedef struct {
  int64_t swlevel;
  int64_t swsize;
  double atol;
  double rtol;
  double c_tol;
  int64_t leafsize;
  int64_t kest;
  int64_t stepsize;
  int32_t verbose;
  int32_t subtree;    /* extension for subtree-per-GPU runs: 1 = the tree is a subtree of a larger one — its root keeps a
                         non-empty boundary whose Schur block is exported (hs_schur_export) instead of being solved */
  int32_t hss;        /* extension: 1 = HSS storage of compressed nodes' Schur complements (see above), 0 = dense */
  int32_t pad0;
  /* Host-supplied Gaussian test matrices of the randomized HSS construction (the parity anchor: the reference draws them
   * from Julia's global RNG, test/rungmres.jl:7).  Column-major sketch_rows × sketch_cols, element type of the
   * factorization; a compressed node whose S has m rows and needs k samples uses Ω[0:m, 0:k] and Ψ[0:m, 0:k] (S·Ω and
   * Sᴴ·Ψ).  sketch_rows must cover the largest compressed boundary and sketch_cols the largest sample count reached
   * (HS_EARG otherwise).  NULL: both are drawn on the device from a counter-based generator seeded with sketch_seed. */
  const void* sketch_omega;
  const void* sketch_psi;
  int64_t sketch_rows, sketch_cols;
  uint64_t sketch_seed;
} hs_opts;

/* Serialized elimination tree, the ragged form of the `.mat` schema that parse_elimtree consumes
 * (nesteddissection.jl:105-148, util/read_problem.jl:14-20).  Node ids and DOF ids use `index_base`;
 * `-1` means "none" for fathers/lsons/rsons regardless of the base (nesteddissection.jl:110,122). */
typedef struct {
  int64_t nnodes;
  const int64_t* fathers;
  const int64_t* lsons;
  const int64_t* rsons;
  const int64_t* inter_ptr; /* nnodes+1, 0-based offsets into inter_idx */
  const int64_t* inter_idx;
  const int64_t* bound_ptr;
  const int64_t* bound_idx;
  int32_t index_base;
} hs_elimtree;

/* What `symfact!` returns, flattened: the pair (nd, nd_loc) of nesteddissection.jl:29-69.
 * Nodes are numbered 0..nnodes-1 in post-order (children before parents, root last); `left/right` are node
 * numbers or -1.  `int/bnd` are global DOF ids, `iloc/bloc` are positions inside the node's own `bnd`
 * (nd_loc.int / nd_loc.bnd), all using `index_base`. */
typedef struct {
  int64_t nnodes;
  const int64_t* left;
  const int64_t* right;
  const int64_t* int_ptr;
  const int64_t* int_idx;
  const int64_t* bnd_ptr;
  const int64_t* bnd_idx;
  const int64_t* iloc_ptr;
  const int64_t* iloc_idx;
  const int64_t* bloc_ptr;
  const int64_t* bloc_idx;
  int32_t index_base;
} hs_tree;

typedef struct hs_ctx hs_ctx;
typedef struct hs_symbolic hs_symbolic; /* host-side result of hs_symfact (owns the arrays an hs_tree views) */
typedef struct hs_fac hs_fac;           /* device-resident FactorNode tree                                 */

/* per-phase counters of the last factor / solve, CUDA-event timed on the context's stream */
typedef struct {
  int64_t nnodes, nlevels, n;
  int64_t max_ni, max_nb;
  double factor_flops;     /* Σ ⅔ni³ + 2ni²nb + 2ni·nb² (×4 complex)                                      */
  double solve_bytes;      /* esz · Σ (ni² + 2 ni nb) per right-hand side                                 */
  double extadd_bytes;     /* 2 · esz · Σ nb²                                                            */
  double front_bytes;      /* device bytes held by the fronts                                            */
  double ms_analyze, ms_h2d, ms_assemble, ms_panel, ms_trsm, ms_gemm, ms_factor_total;
  double ms_solve_fwd, ms_solve_bwd, ms_solve_total;
  int64_t launches_factor, launches_solve;
  int64_t singular_front, singular_col; /* -1 when the factorization is non-singular                    */
  int64_t maxrank;
  double gemm_flops;       /* flops issued by the Schur/trailing-update GEMM launches (2·m²·k per step, ×4 complex) */
  int64_t gemm_launches, panel_launches;
  double ms_extend_add;    /* child Schur blocks → parent fronts (part of ms_assemble) */
  double ms_small;         /* fused small-front kernel (levels whose fronts fit in registers) */
  double ms_solve_prep;    /* in-place inversion of the diagonal blocks of L11/U11 (part of the factor time) */
  double ms_compress;      /* compressed fronts: pivoted QR of A_bi / A_ib, thin fronts, Schur complement      */
  double lowrank_bytes;    /* device bytes of the thin fronts + low-rank factors of compressed fronts (their dense
                              slots are transient and counted once, as two arenas, in front_bytes)            */
  double ms_hss;           /* randomized HSS construction of the Schur complements (sketches, IDs, couplings, expansion) */
  double hss_bytes;        /* device bytes of the stored HSS generators                                          */
  int64_t hss_maxrank;     /* max over compressed nodes of hssrank(S) (factornode.jl:53)                          */
  int64_t hss_rounds;      /* adaptive rounds taken in total (1 per compressed level when no rank saturated)      */
  int64_t hss_nodes;       /* HSS tree nodes over all compressed fronts                                           */
  double sketch_flops;     /* flops of the sketch GEMMs S·Ω, Sᴴ·Ψ (matrix-free: Abb·X − Z·(Ri·X))                 */
  double gemm_flops_big;   /* part of gemm_flops issued by the K = outer-block trailing updates (the rest are the K = panel-width
                              in-block updates, launches of a few CTAs bounded by their prologue / epilogue)       */
  double ms_gemm_big;      /* their share of ms_gemm (HS_PROFILE)                                                 */
} hs_stats_t;

typedef enum { HS_GET_D = 0, HS_GET_S = 1, HS_GET_L = 2, HS_GET_R = 3, HS_GET_FRONT = 4, HS_GET_PIV = 5 } hs_which;

int32_t hs_version(void);
const char* hs_last_error(void);

/* context: one per GPU / host thread.  `stream` may be NULL (library creates its own). */
int32_t hs_create(hs_ctx** out, int32_t device);
int32_t hs_set_stream(hs_ctx* ctx, void* cuda_stream);
int32_t hs_destroy(hs_ctx* ctx);
/* per-phase CUDA-event timing (serialises the stream; off by default, HS_PROFILE=1 turns it on at hs_create) */
int32_t hs_set_profile(hs_ctx* ctx, int32_t on);
/* number of kernels this context has launched so far */
int32_t hs_launch_count(hs_ctx* ctx, int64_t* count);
/* 1 when the shared library was built with CUDA kernels and a device is usable; never falls back to CPU */
int32_t hs_device_count(void);

/* ---- symbolic phase, host only -------------------------------------------------------------------
 * replaces parse_elimtree + symfact! + postorder + permuted!(nd, invperm(perm))
 * (nesteddissection.jl:105-148, :29-69, :73-79, :82-88; driver order test/rungmres.jl:15-19).
 * `apply_postorder` != 0 renumbers DOFs by the post-order permutation exactly as rungmres.jl:17-19 does;
 * the caller must then factor permute(A, perm, perm).  perm is reported with `index_base`. */
/* ---- elimination tree for a general sparsity pattern ----------------------------------
 * The reference loads `elim_tree` from the problem file (util/read_problem.jl:14-20) and has no ordering code.  This
 * builds the same schema by recursive graph bisection of the pattern of A + Aᵀ (METIS, the static library of the CUDA
 * toolkit) until a part holds at most `nmax` DOFs.  A is CSC (only colptr / rowval are read); `flags` as for
 * hs_factor (HS_CSC_ZERO_BASED, HS_CSC_INT32); node and DOF ids of the result use `index_base`. */
typedef struct hs_ordering hs_ordering;
int32_t hs_nd_create(int64_t n, const void* colptr, const void* rowval, int32_t flags, int32_t index_base, int64_t nmax,
                     hs_ordering** out);
int32_t hs_nd_elimtree(const hs_ordering* o, hs_elimtree* et_out);   /* views into `o` */
int32_t hs_nd_free(hs_ordering* o);

int32_t hs_symfact(const hs_elimtree* et, int32_t apply_postorder, hs_symbolic** out);
int32_t hs_symbolic_tree(const hs_symbolic* s, hs_tree* tree_out);      /* views into `s` */
int32_t hs_symbolic_perm(const hs_symbolic* s, const int64_t** perm, int64_t* n);
int32_t hs_symbolic_depth(const hs_symbolic* s, int64_t* depth);
int32_t hs_symbolic_free(hs_symbolic* s);

/* ---- numeric factorization -----------------------------------------------------------------------
 * replaces factor(A, nd, nd_loc, opts; kw...) → FactorNode   (factorization.jl:5-11 and everything below it:
 * _factor :14-27, _factor_leaf :30-42, _factor_branch :62-75, _assemble_blocks :115-123,
 * blockfactor/blockldiv/blockrdiv blockmatrix.jl:115-187).
 * A is CSC (SparseMatrixCSC: colptr n+1, rowval nnz, nzval nnz) with `index_base` of the tree.
 * `flags`: HS_ON_DEVICE — colptr/rowval/nzval are device pointers already resident in HBM (0-based int64);
 *          HS_CSC_ZERO_BASED — the host CSC arrays are 0-based whatever the tree's base is (SciPy callers);
 *          HS_CSC_INT32 — colptr/rowval point to int32 arrays (SciPy's default index type). */
#define HS_ON_DEVICE 1
#define HS_CSC_ZERO_BASED 2
#define HS_CSC_INT32 4
int32_t hs_factor(hs_ctx* ctx, hs_dtype dtype, int64_t n, const int64_t* colptr, const int64_t* rowval,
                  const void* nzval, const hs_tree* tree, const hs_opts* opts, int32_t flags, hs_fac** out);
/* same plan and sparsity, new values (the numeric part of `factor` alone); nzval == NULL re-uses the stored values */
int32_t hs_refactor(hs_fac* fac, const void* nzval, int32_t on_device);
/* hs_factor without the numeric phase (plan + upload only); follow with hs_schur_import / hs_refactor */
int32_t hs_analyze(hs_ctx* ctx, hs_dtype dtype, int64_t n, const int64_t* colptr, const int64_t* rowval,
                   const void* nzval, const hs_tree* tree, const hs_opts* opts, int32_t flags, hs_fac** out);

/* ---- subtree-per-GPU plumbing -------------------------------------------------------
 * Disjoint bottom subtrees are factored one per GPU with opts.subtree = 1; each exports the Schur complement of its
 * root (nb×nb, rows/cols in the node's `bnd` order, column-major with leading dimension ld) into a device buffer that
 * the host moves over NCCL.  The upper part of the tree is a tree whose leaves are those roots with an EMPTY `int`:
 * hs_schur_import registers the device buffer such a leaf's front is copied from (instead of gathering A) by the
 * next hs_refactor.  Child boundaries are disjoint, so the exchange is a concatenation, never a reduction. */
int32_t hs_schur_export(hs_fac* fac, int64_t node, void* dst_dev, int64_t ld);
int32_t hs_schur_import(hs_fac* fac, int64_t node, const void* src_dev, int64_t ld);
/* one half of ldiv! in place on a device-resident n×nrhs block: which = 1 forward (post-order), 2 backward
 * (pre-order), 3 both.  With opts.subtree the root's boundary rows are updated (forward) / consumed (backward). */
int32_t hs_solve_sweep(hs_fac* fac, int64_t nrhs, void* x_dev, int64_t ldx, int32_t which);
int32_t hs_factor_free(hs_fac* fac);

/* replaces ldiv!(C, F, B) for vectors and matrices (factornode.jl:62-74: _lsolve! :77, _dsolve! :89,
 * root Schur solve :72, _rsolve! :83).  B and X may alias.  on_device != 0: B/X are device pointers. */
int32_t hs_solve(hs_fac* fac, int64_t nrhs, const void* B, int64_t ldb, void* X, int64_t ldx, int32_t on_device);

/* FactorNode field access (factornode.jl:8-22), node numbering of hs_tree.  `dims[2]` receives rows, cols;
 * call with out == NULL to query dims.  D/S/L/R are returned exactly as the reference defines them
 * (D = the pivot block A_ii, L = A_bi·A_ii⁻¹, R = A_ii⁻¹·A_ib, S = Schur complement permuted by
 * [int_loc; bnd_loc]); HS_GET_FRONT returns the raw partially factored front, HS_GET_PIV its pivots (int64). */
int32_t hs_node_get(hs_fac* fac, int64_t node, hs_which which, void* out, int64_t* dims);
int32_t hs_maxrank(hs_fac* fac, int64_t* rank);              /* factornode.jl:49-57: max over nodes of hssrank(S), rank(L), rank(R) */
/* rank(F.L), rank(F.R) of one node (LowRankMatrix, factorization.jl:173,179); 0, 0 for an uncompressed node whose
 * L and R are dense.  For a compressed node hs_node_get returns the dense products L = U·Vᴴ, R = U·Vᴴ. */
int32_t hs_node_rank(hs_fac* fac, int64_t node, int64_t* rank_l, int64_t* rank_r);
/* HSS form of a compressed node's Schur complement (`F.S::HssMatrix`, factorization.jl:110-111; fields of
 * HssMatrices.jl's HssMatrix: leaf D, U, V; branch B12, B21, R1/R2 stacked as R, W1/W2 stacked as W).  HSS tree nodes of
 * one front are numbered in pre-order (0 = root, then the whole A11 subtree, then A22).
 *   hs_hss_info(fac, node, &nhss, info): nhss = number of HSS tree nodes (0: S is dense); info (may be NULL, else
 *     8·nhss int64): per HSS node {lo, hi, left, right, rank_u, rank_v, parent, is_leaf}; rows [lo, hi) of S[perm,perm].
 *   hs_hss_get(fac, node, hnode, which, out, dims): one generator, column-major; out == NULL queries dims. */
typedef enum { HS_HSS_D = 0, HS_HSS_U = 1, HS_HSS_V = 2, HS_HSS_B12 = 3, HS_HSS_B21 = 4, HS_HSS_R = 5, HS_HSS_W = 6 } hs_hss_which;
int32_t hs_hss_info(hs_fac* fac, int64_t node, int64_t* nhss, int64_t* info);
int32_t hs_hss_get(hs_fac* fac, int64_t node, int64_t hnode, hs_hss_which which, void* out, int64_t* dims);
int32_t hs_stats(hs_fac* fac, hs_stats_t* out);
int32_t hs_resolved_swlevel(hs_fac* fac, int64_t* swlevel);  /* factorization.jl:8 */
/* Device-resident copy of the matrix a factorization holds (0-based int64 colptr / rowval, nzval of the factorization's
 * dtype): lets further hs_factor / hs_analyze calls on the same matrix (the upper fronts of the subtree-per-GPU mapping)
 * pass HS_ON_DEVICE | HS_CSC_ZERO_BASED instead of uploading the CSC arrays again.  Valid until hs_factor_free(fac). */
/* y = A·x on the device (x, y device vectors of the factorization's dtype, asynchronous on the context's stream) with the
 * matrix `fac` holds — the mat-vec of a Krylov loop driven from the host side (replicated GMRES of the multi-GPU path). */
int32_t hs_spmv(hs_fac* fac, const void* x, void* y);
/* wrapping sum and XOR of the 64-bit words of the matrix values held in HBM: lets a host binding check cheaply that a
 * matrix it is handed still equals the resident copy before skipping the upload (hs_gmres with colptr == NULL) */
int32_t hs_matrix_checksum(hs_fac* fac, uint64_t* sum_out, uint64_t* xor_out);
int32_t hs_matrix_device(hs_fac* fac, const int64_t** colptr, const int64_t** rowval, const void** nzval, int64_t* nnz);

/* ---- GMRES with the factorization as right preconditioner (test/rungmres.jl:47-48) ---------------
 * device-resident restarted GMRES (modified Gram-Schmidt, Givens), x0 = 0, stops when the running residual
 * estimate ≤ reltol·‖b‖ or after maxiter Arnoldi steps.  `fac` may be NULL (no preconditioner).
 * resnorm must hold maxiter doubles; niter receives the number of Arnoldi steps taken.
 * colptr == NULL: iterate on the matrix held by `fac` (the one it factored).  on_device != 0: b and x are device
 * pointers. */
int32_t hs_gmres(hs_ctx* ctx, hs_dtype dtype, int64_t n, const int64_t* colptr, const int64_t* rowval,
                 const void* nzval, int32_t index_base, hs_fac* fac, const void* b, void* x, double reltol,
                 int64_t restart, int64_t maxiter, double* resnorm, int64_t* niter, int32_t* converged,
                 int32_t on_device);

#ifdef __cplusplus
}
#endif
#endif /* HSOLVE_CUDA_H */

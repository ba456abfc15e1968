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

#include <memory>

// int table whose pages are first touched by the threads that fill it (std::vector would zero-fill it serially)
struct IntBuf {
  std::unique_ptr<int[]> p;
  size_t n = 0, cap = 0;
  void reserve(size_t c) { if (c > cap || !p) { p.reset(new int[std::max<size_t>(c, 1)]); cap = c; } n = 0; }
  void set_size(size_t m) { if (m > cap) throw hs_error(HS_ECUDA, "plan table overflow"); n = m; }
  void resize(size_t m, int fill) { const size_t o = n; set_size(m); for (size_t i = o; i < m; ++i) p[i] = fill; }
  void push_back(int v) { set_size(n + 1); p[n - 1] = v; }
  int* data() { return p.get(); }
  size_t size() const { return n; }
  int& operator[](size_t i) { return p[i]; }
};

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct hs_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t aux_stream = nullptr;  // look-ahead: the big trailing update overlaps the next block's panels
  cudaEvent_t ev_b = nullptr, ev_c2 = nullptr;
  // boundary rows (the rows below the ones pivots are taken from) are solved and updated on their own stream: the chain
  // panel → row interchanges + U row panel → update of the pivot rows → next panel does not wait for them
  cudaStream_t below_stream = nullptr;
  cudaEvent_t ev_pan = nullptr, ev_urow = nullptr, ev_below = nullptr;
  cudaStream_t prep_stream = nullptr;  // solve preparation of a finished level runs beside the next levels' factorization
  cudaEvent_t ev_p0 = nullptr, ev_p1 = nullptr;
  bool prep_pending = false;
  int lookahead_max_fronts = 32;
  int max_cluster = 8;
  bool profile = false;
  long long launches = 0;
  void* gm_buf = nullptr;   // GMRES workspace (Krylov basis + work vectors), grow-only, reused across hs_gmres calls
  size_t gm_bytes = 0;
  int outer_block = 256;  // NB of the two-level blocked LU (HS_OUTER_BLOCK)
  IntBuf sc_gidx, sc_cmap;  // plan-build scratch (row tables before their upload), grow-only
};

// ------------------------------------------------------------------------------------------------
// factorization object
// ------------------------------------------------------------------------------------------------
struct Level {
  int f0 = 0, f1 = 0;        // front range
  int max_n = 0, max_ni = 0, max_nb = 0;
  long long ioff0 = 0, ioff1 = 0;
  long long poff0 = 0, poff1 = 0;
  long long tpoff0 = 0, tpoff1 = 0;  // dense slots of the level's compressed fronts inside a transient arena
  bool pseudo = false;
  std::vector<int> ni_sorted;  // ni of the fronts in this level (descending)
  int max_prow = 0;            // tallest diagonal block pivots are searched in (max ni, or max(split, ni − split) of split fronts)
  bool has_split = false;
  int fm = 0;                  // assembly levels: fronts [fm, f1) are compressed (fm = f1: none)
  int thin = -1;               // factor/solve levels: index into hs_fac::clevels when this is a level of thin fronts
};

// ---- compressed path (factorization.jl:78-112): low-rank Gauss transforms ------------------------------------
// A compressed front keeps its dense slot in the pool for the assembly and for its Schur block S, but what is
// factored and kept for the solve is the THIN bordered front
//        [ Aii   Qi ]  ni            Abi ≈ Qb·Rb   (nb×r1 · r1×ni)      L = Qb·(Rb·Aii⁻¹)
//        [ Rb    0  ]  r             Aib ≈ Qi·Ri   (ni×r2 · r2×nb)      R = (Aii⁻¹·Qi)·Ri
// whose partial LU yields Rb·U11⁻¹ and L11⁻¹·P·Qi: the solve runs through the ordinary front kernels on ni + r rows,
// the r border rows living in "virtual" slots appended to the solution vector, plus one thin product with Qb
// (forward) and one with Ri (backward).
struct CompFront {
  int fi = 0;        // dense front id; the thin descriptor is fronts[nfr + fi]
  int ni = 0, nb = 0;
  int rcap = 0;      // min(ni, nb): largest rank a pivoted QR of Abi / Aib can reveal
  int vcap = 0;      // virtual slots reserved for the thin front's border rows: rcap, or ni + nb for a front whose L, R start
                     // from its HSS children's generators (their ranks add up without recompression, factorization.jl:184-209)
  long long yoff = 0, zoff = 0;   // Y = Aii⁻¹·Qi (ni×r2) and Z = Abi·Y (nb×r2) in the level's side buffer, offsets from pool
  int r1 = 0, r2 = 0, r = 0;
  long long voff = 0;                   // first virtual slot (x index n + voff)
  long long qb = 0, vit = 0, ri = 0;    // element offsets from `pool` of Qb (nb×r1), Riᵀ (nb×r2), Ri (r2×nb)
  int qb_ld = 0, ri_ld = 0;
  int hss = -1;      // index into hs_fac::hss when this front's Schur complement is stored as an HSS matrix
  int rx1 = 0, rx2 = 0;  // HSS-children Gauss transforms: ranks of the sketched sparse couplings appended to L / R
  int hchild = 0;    // 1: both children are HSS with the split at their int/bnd boundary → generator-concatenating transforms
  int cl = -1, cr = -1;           // hchild: indices of the two children in hs_fac::comp
  int ni_l = 0, nb_l = 0;         // rows of int / bnd that come from the left child
  long long x1 = 0, x2 = 0;       // hchild: anti-diagonal copies of Abi (nb×ni) / Aib (ni×nb) in CompLevel::xws, offsets from pool
  int ldx1 = 0, ldx2 = 0;
};
struct IdRun {       // one truncated column-pivoted QR (pivoted Cholesky of the Gram matrix), device + host
  long long moff;    // pool offset of M(0,0);  M is m × ncol with leading dimension ld
  int m, ncol, ld;
  int rcap, ldr;     // at most rcap steps; R is ldr × ncol in the workspace at rws
  int ip;            // ints[ip .. ip+rcap): pivots
  long long rws;
  long long st;      // state (doubles): d[2][ncol], then r11
  int slot;          // ints[1 + slot]: rank (−1 while running)
  int pad;
};
struct LrDesc {      // device image of one compressed front once its ranks are known
  int fi, thin;
  int ni, nb;
  int r1, r2, r, pad;
  long long voff;
  long long qb, vit, ri;
  int qb_ld, ri_ld;
  // HSS-children fronts: the first off1 (off2) columns of Qb (Qi) / rows of Rb (Ri) come from the children's generators,
  // the pivoted QR of the sparse couplings contributes r1x (r2x) more; dense-children fronts: off = 0, r?x = r?
  int off1, off2, r1x, r2x;
  long long y, z;      // Y = Aii⁻¹·Qi, Z = Abi·Y
  long long x1, x2;    // copies of the anti-diagonal (sparse coupling) blocks of Abi / Aib the pivoted QR runs on
  int ldx1, ldx2, hchild, ni_l, nb_l, pad2;
};
struct CopyDesc {      // one block of a batched copy (hs_hss.cu: k_copy_desc)
  long long src, dst;
  int lds, ldd, rows, cols;
  int gap_at, gap_skip;   // logical source row i ≥ gap_at lives at physical row i + gap_skip
  int mode;               // 0: dst[i,j] = src[i,j]   1: dst[j,i] = conj(src[i,j])   2: dst[j,i] = src[i,j]
  int pad;
};
struct CompLevel {
  int li = 0;             // assembly level
  int c0 = 0, c1 = 0;     // range in hs_fac::comp; runs 2·c (Abi) and 2·c + 1 (Aib)
  int flevel = 0;         // index of the thin level in hs_fac::flevels
  void* side = nullptr;   // thin fronts + Qb + Ri of this level (sized once the ranks are known)
  size_t side_bytes = 0;
  void* hss_store = nullptr;   // HSS generators of this level's Schur complements (grow-only across refactorizations)
  size_t hss_store_bytes = 0;
  void* xws = nullptr;         // workspace of the HSS-children Gauss transforms (copies of the sparse couplings)
  size_t xws_bytes = 0;
};

// ---- HSS storage of the Schur complements (factorization.jl:102-110; HssMatrices.jl's HssMatrix) ----------------------
struct HssTreeNode {       // cluster-tree node of one front, numbered in pre-order inside the front
  int lo = 0, hi = 0;      // rows [lo, hi) of S[perm,perm]
  int left = -1, right = -1, parent = -1;
  int height = 0, isright = 0;
};
struct HssStored {         // generators of one HSS node; element offsets from the start of the level's store (add HssFront::sbase)
  int r0 = 0, r1 = 0;      // rank of the row basis (U / [R1;R2]) and of the column basis (V / [W1;W2])
  long long D = -1, U = -1, VH = -1;   // leaf: D m×m, U m×r0, Vᴴ r1×m
  long long R = -1, WH = -1;           // non-root branch: R (r0a+r0b)×r0, Wᴴ r1×(r1a+r1b)
  long long B12 = -1, B21 = -1;        // branch: B12 r0a×r1b, B21 r0b×r1a
  int ldD = 0, ldU = 0, ldVH = 0, ldR = 0, ldWH = 0, ldB12 = 0, ldB21 = 0;
};
struct HssFront {
  int comp = -1;           // index into hs_fac::comp
  int m = 0, n1 = 0;       // rows of S[perm,perm] and the forced first split (factorization.jl:109)
  std::vector<HssTreeNode> tree;
  std::vector<HssStored> st;
  std::vector<int> perm;   // [int_loc; bnd_loc], 0-based positions in the node's bnd
  int perm_off = 0;        // offset of `perm` inside hs_fac::d_hperm
  int k = 0, rounds = 0;   // sample count of the last round, adaptive rounds taken
  bool done = false;
  int hssrank = 0;
  long long sbase = 0;     // start of the level's store as an element offset from `pool` (may be negative)
  // what the parent's Gauss transforms take from this HSS matrix (:129-137), element offsets into the store (add sbase):
  // ta = Û(A11)·B12 (n1×rb1), tb = Û(A22)·B21 ((m−n1)×ra1), vha = V̂(A11)ᴴ (ra1×n1), vhb = V̂(A22)ᴴ (rb1×(m−n1))
  long long ta = -1, tb = -1, vha = -1, vhb = -1;
  int ld_ta = 0, ld_tb = 0, ld_vha = 0, ld_vhb = 0;
  int ra1 = 0, rb1 = 0;
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
  std::vector<int64_t> iloc_ptr, bloc_ptr;
  IntBuf iloc_idx, bloc_idx;   // positions inside the node's bnd (0-based)
  std::vector<int> node_ni, node_nb, node2front;
  std::vector<Front> fronts;
  std::vector<Level> levels;  // assembly levels: deepest first, root last (+ pseudo front for a non-empty root boundary)
  std::vector<Level> flevels; // factor / solve levels: per assembly level its dense fronts, then its thin fronts
  std::vector<CompFront> comp;
  std::vector<CompLevel> clevels;
  std::vector<HssFront> hss;   // HSS Schur complements (opts.hss); CompFront::hss indexes it
  void* d_sk = nullptr;        // sketch matrices Ω, Ψ on the device (sk_rows × sk_cols each, Ψ behind Ω)
  long long sk_rows = 0, sk_cols = 0;
  int* d_hperm = nullptr;      // perm of every HSS front, concatenated
  // device scratch of the HSS construction is recycled through this free list (cudaMalloc / cudaFree per adaptive round
  // and per level cost more than the kernels they serve); released with the factorization
  std::vector<std::pair<void*, size_t>> dev_cache;
  int nfr = 0;                // number of dense front descriptors; thin descriptors follow at nfr + fi
  bool transient_schur = true; // dense slots of compressed fronts are recycled two levels up
  long long nvirt = 0;        // virtual slots appended to the solution vector
  long long xld = 0;          // leading dimension of the internal solution buffer: n + nvirt
  std::vector<IdRun> runs;    // two per compressed front
  void* d_cws = nullptr;      // workspace of a level: R factors, later Aii⁻¹Qi and Abi·Aii⁻¹Qi
  double* d_cstate = nullptr; // per run: running column norms (2×ncol), r11
  int* d_cint = nullptr;      // [0] runs finished, [1+slot] ranks, pivots
  IdRun* d_runs = nullptr;
  LrDesc* d_lr = nullptr;     // one per compressed front (index = position in `comp`)
  void* d_gd = nullptr;       // GemmDesc staging
  size_t gd_cap = 0;
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
  long long *d_csr_ptr = nullptr, *d_csr_col = nullptr, *d_csr_src = nullptr;  // d_csr_src: CSC position of every CSR entry
  void* d_csr_val = nullptr;
  bool csr_stale = true;      // values of the CSR image are older than d_nzval (hs_refactor with new values)
  int64_t rhs_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  hs_stats_t stats{};

  ~hs_fac() {
    cudaFree(pool); cudaFree(d_fronts); cudaFree(d_gidx); cudaFree(d_ipiv); cudaFree(d_rperm); cudaFree(d_cmap);
    cudaFree(d_own); cudaFree(d_pos); cudaFree(d_info); cudaFree(d_colptr); cudaFree(d_rowval); cudaFree(d_nzval);
    cudaFree(d_x); cudaFree(d_work); cudaFree(d_csr_ptr); cudaFree(d_csr_col); cudaFree(d_csr_val); cudaFree(d_csr_src);
    cudaFree(d_cws); cudaFree(d_cstate); cudaFree(d_cint); cudaFree(d_runs); cudaFree(d_lr); cudaFree(d_gd);
    for (auto& c : clevels) { cudaFree(c.side); cudaFree(c.hss_store); cudaFree(c.xws); }
    cudaFree(d_sk); cudaFree(d_hperm);
    for (auto& b : dev_cache) cudaFree(b.first);
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
void hs_trsm_rows_f64(hs_fac* f, int W, int f0, int nact, int j0, int max_rows, cudaStream_t st);
void hs_trsm_rows_c64(hs_fac* f, int W, int f0, int nact, int j0, int max_rows, cudaStream_t st);
inline void hs_trsm_rows(hs_fac* f, int W, int f0, int nact, int j0, int max_rows, cudaStream_t st) {
  if (f->dtype == HS_F64) hs_trsm_rows_f64(f, W, f0, nact, j0, max_rows, st); else hs_trsm_rows_c64(f, W, f0, nact, j0, max_rows, st);
}
inline int hs_panel_width(const hs_fac* f, int max_n, int nfronts) {
  return f->dtype == HS_F64 ? hs_panel_width_f64(f, max_n, nfronts) : hs_panel_width_c64(f, max_n, nfronts);
}
inline void hs_panel_launch(hs_fac* f, int W, int f0, int nact, int j0, int m, cudaStream_t st) {
  if (f->dtype == HS_F64) hs_panel_launch_f64(f, W, f0, nact, j0, m, st); else hs_panel_launch_c64(f, W, f0, nact, j0, m, st);
}

// hs_solve.cu
void hs_solve_setup();
int hs_solve_block(hs_dtype dt);
void hs_solve_prep(hs_fac* f, const Level& L, cudaStream_t st);
void hs_solve_run(hs_fac* f, int64_t nrhs, void* x, int which);  // which: 1 forward, 2 backward, 3 both

// hs_small.cu
int hs_small_max_n(hs_dtype dt);
void hs_small_factor(hs_fac* f, const Level& L);

// hs_compress.cu
void hs_comp_setup();
void hs_comp_plan(hs_fac* f);                        // workspaces + device descriptors, after build_plan
void hs_comp_prepare(hs_fac* f, CompLevel& C);       // IDs of Abi / Aib, ranks, thin fronts (before their LU)
void hs_comp_schur(hs_fac* f, CompLevel& C);         // S = Abb − (Abi·Aii⁻¹Qi)·Ri (after the thin LU)
void hs_comp_solve(hs_fac* f, const CompLevel& C, int64_t nrhs, void* x, bool fwd);
struct GemmDesc;
void hs_gen_gemm(hs_fac* f, const GemmDesc* d_items, int nitems, int maxM, int maxN);  // batched C ∓= A·B on the DMMA kernel

// hs_hss.cu
void hs_hss_setup();
void hs_hss_plan(hs_fac* f);                         // cluster trees of the compressed fronts, perm tables, sketch matrices
void hs_hss_build(hs_fac* f, CompLevel& C);          // randomized adaptive HSS construction + expansion into the dense slots
void hs_hss_dense(hs_fac* f, int hi, void* out_host);
void hs_hss_copy(hs_fac* f, const std::vector<CopyDesc>& blocks);   // batched block copies on the context's stream (synchronous)  // dense S[perm,perm] a stored HSS matrix represents (m×m, column-major)

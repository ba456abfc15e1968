// Nested-dissection elimination trees for general sparse matrices.
//
// The reference does not compute orderings: `read_problem` (util/read_problem.jl:5-25) loads an `elim_tree` that came
// with the matrix.  This produces that same schema (fathers / lsons / rsons / inter / bound, what `parse_elimtree`
// consumes, nesteddissection.jl:105-148) for an arbitrary sparsity pattern, with the invariants `_symfact!` and
// `_factor_*` rely on (nesteddissection.jl:42-65, factorization.jl:30-42,62-75):
//   * the vertex sets of the leaves partition the DOFs; leaf.inter = vertices all of whose neighbours are in the leaf,
//     leaf.bound = the rest;
//   * branch.inter ∪ branch.bound = left.bound ∪ right.bound (disjoint), inter = the vertices without a neighbour
//     outside the branch's vertex set; the root's bound is empty.
// The tree is a recursive graph BISECTION (edge cuts, not vertex separators — that is the schema's model): METIS
// `METIS_PartGraphRecursive` with two parts on the induced subgraph until a part has at most `nmax` vertices.
// METIS is the static library that ships with the CUDA toolkit (libmetis_static.a, 64-bit idx_t, 32-bit real_t).
#include "hs_internal.h"

#include <algorithm>
#include <cstring>
#include <numeric>

extern "C" {
int METIS_SetDefaultOptions(int64_t* options);
int METIS_PartGraphRecursive(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* vsize,
                             int64_t* adjwgt, int64_t* nparts, float* tpwgts, float* ubvec, int64_t* options,
                             int64_t* objval, int64_t* part);
}

struct hs_ordering {
  int64_t n = 0, nnodes = 0;
  int32_t index_base = 0;
  std::vector<int64_t> fathers, lsons, rsons, inter_ptr, inter_idx, bound_ptr, bound_idx;
};

namespace {

struct TNode {
  int64_t v0 = 0, v1 = 0;  // range in the vertex array (the vertices of a subtree are kept contiguous)
  int64_t left = -1, right = -1, father = -1;
};

}  // namespace

extern "C" int32_t hs_nd_create(int64_t n, const void* colptr, const void* rowval, int32_t flags, int32_t index_base,
                                int64_t nmax, hs_ordering** out) {
  HS_TRY_BEGIN
  if (!colptr || !rowval || !out) return hs_fail(HS_EARG, "hs_nd_create: null argument");
  if (n <= 0 || nmax < 1) return hs_fail(HS_EARG, "hs_nd_create: n and nmax must be positive");
  const bool idx32 = flags & HS_CSC_INT32;
  const int64_t base = (flags & HS_CSC_ZERO_BASED) ? 0 : index_base;
  auto cp = [&](int64_t j) -> int64_t { return (idx32 ? (int64_t)((const int32_t*)colptr)[j] : ((const int64_t*)colptr)[j]) - base; };
  auto rv = [&](int64_t p) -> int64_t { return (idx32 ? (int64_t)((const int32_t*)rowval)[p] : ((const int64_t*)rowval)[p]) - base; };
  const int64_t nnz = cp(n);
  if (cp(0) != 0 || nnz < 0) return hs_fail(HS_EARG, "hs_nd_create: malformed colptr");
  // adjacency of the pattern of A + Aᵀ without the diagonal
  std::vector<int64_t> deg(n + 1, 0);
  for (int64_t j = 0; j < n; ++j)
    for (int64_t p = cp(j); p < cp(j + 1); ++p) {
      const int64_t i = rv(p);
      if (i < 0 || i >= n) return hs_fail(HS_EARG, "hs_nd_create: row index out of range");
      if (i != j) { ++deg[i + 1]; ++deg[j + 1]; }
    }
  std::vector<int64_t> xadj(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) xadj[i + 1] = xadj[i] + deg[i + 1];
  std::vector<int64_t> adj(xadj[n]), pos(xadj.begin(), xadj.end() - 1);
  for (int64_t j = 0; j < n; ++j)
    for (int64_t p = cp(j); p < cp(j + 1); ++p) {
      const int64_t i = rv(p);
      if (i != j) { adj[pos[i]++] = j; adj[pos[j]++] = i; }
    }
  {  // sort + unique per vertex (an entry present in both triangles was added twice)
    std::vector<int64_t> nx(n + 1, 0), na;
    na.reserve(adj.size());
    for (int64_t i = 0; i < n; ++i) {
      std::sort(adj.begin() + xadj[i], adj.begin() + xadj[i + 1]);
      auto e = std::unique(adj.begin() + xadj[i], adj.begin() + xadj[i + 1]);
      na.insert(na.end(), adj.begin() + xadj[i], e);
      nx[i + 1] = (int64_t)na.size();
    }
    xadj.swap(nx);
    adj.swap(na);
  }
  // recursive bisection; `verts` is permuted in place so that every tree node owns a contiguous range
  std::vector<int64_t> verts(n);
  std::iota(verts.begin(), verts.end(), 0);
  std::vector<TNode> tree(1);
  tree[0].v0 = 0; tree[0].v1 = n;
  std::vector<int64_t> loc(n, -1), sx, sa, part, tmp;
  for (size_t q = 0; q < tree.size(); ++q) {  // breadth-first: ids in creation order, as the grid generator numbers them
    const int64_t v0 = tree[q].v0, v1 = tree[q].v1, m = v1 - v0;
    if (m <= nmax || m < 2) continue;
    // induced subgraph in local numbering
    for (int64_t k = 0; k < m; ++k) loc[verts[v0 + k]] = k;
    sx.assign(m + 1, 0);
    sa.clear();
    for (int64_t k = 0; k < m; ++k) {
      const int64_t v = verts[v0 + k];
      for (int64_t p = xadj[v]; p < xadj[v + 1]; ++p)
        if (loc[adj[p]] >= 0) sa.push_back(loc[adj[p]]);
      sx[k + 1] = (int64_t)sa.size();
    }
    part.assign(m, 0);
    bool ok = false;
    if (!sa.empty()) {
      int64_t nv = m, ncon = 1, nparts = 2, objval = 0, options[40];
      METIS_SetDefaultOptions(options);
      options[8] = 4321;  // METIS_OPTION_SEED: reproducible trees
      const int rc = METIS_PartGraphRecursive(&nv, &ncon, sx.data(), sa.data(), nullptr, nullptr, nullptr, &nparts, nullptr,
                                              nullptr, options, &objval, part.data());
      int64_t c1 = 0;
      for (int64_t k = 0; k < m; ++k) c1 += part[k] == 1;
      ok = rc == 1 && c1 > 0 && c1 < m;
    }
    if (!ok)  // no edges, or a degenerate answer: halve by position
      for (int64_t k = 0; k < m; ++k) part[k] = k >= m / 2;
    for (int64_t k = 0; k < m; ++k) loc[verts[v0 + k]] = -1;
    tmp.assign(verts.begin() + v0, verts.begin() + v1);
    int64_t w = v0;
    for (int64_t k = 0; k < m; ++k) if (part[k] == 0) verts[w++] = tmp[k];
    const int64_t mid = w;
    for (int64_t k = 0; k < m; ++k) if (part[k] != 0) verts[w++] = tmp[k];
    TNode l, r;
    l.v0 = v0; l.v1 = mid; l.father = (int64_t)q;
    r.v0 = mid; r.v1 = v1; r.father = (int64_t)q;
    tree[q].left = (int64_t)tree.size();
    tree.push_back(l);
    tree[q].right = (int64_t)tree.size();
    tree.push_back(r);
  }
  const int64_t nn = (int64_t)tree.size();
  // position of every vertex in `verts`: v belongs to node q iff v0 ≤ where[v] < v1
  std::vector<int64_t> where(n);
  for (int64_t k = 0; k < n; ++k) where[verts[k]] = k;
  std::vector<std::vector<int64_t>> inter(nn), bound(nn);
  for (int64_t q = nn - 1; q >= 0; --q) {  // children were created after their father
    const TNode& t = tree[q];
    auto outside = [&](int64_t v) {
      for (int64_t p = xadj[v]; p < xadj[v + 1]; ++p) {
        const int64_t w2 = where[adj[p]];
        if (w2 < t.v0 || w2 >= t.v1) return true;
      }
      return false;
    };
    std::vector<int64_t> cand;
    if (t.left < 0) cand.assign(verts.begin() + t.v0, verts.begin() + t.v1);
    else {
      cand = bound[t.left];
      cand.insert(cand.end(), bound[t.right].begin(), bound[t.right].end());
    }
    std::sort(cand.begin(), cand.end());
    for (int64_t v : cand) (outside(v) ? bound[q] : inter[q]).push_back(v);
  }
  std::unique_ptr<hs_ordering> o(new hs_ordering());
  o->n = n; o->nnodes = nn; o->index_base = index_base;
  o->fathers.resize(nn); o->lsons.resize(nn); o->rsons.resize(nn);
  o->inter_ptr.assign(nn + 1, 0); o->bound_ptr.assign(nn + 1, 0);
  for (int64_t q = 0; q < nn; ++q) {
    o->fathers[q] = tree[q].father < 0 ? -1 : tree[q].father + index_base;
    o->lsons[q] = tree[q].left < 0 ? -1 : tree[q].left + index_base;
    o->rsons[q] = tree[q].right < 0 ? -1 : tree[q].right + index_base;
    o->inter_ptr[q + 1] = o->inter_ptr[q] + (int64_t)inter[q].size();
    o->bound_ptr[q + 1] = o->bound_ptr[q] + (int64_t)bound[q].size();
  }
  o->inter_idx.reserve(o->inter_ptr[nn]); o->bound_idx.reserve(o->bound_ptr[nn]);
  for (int64_t q = 0; q < nn; ++q) {
    for (int64_t v : inter[q]) o->inter_idx.push_back(v + index_base);
    for (int64_t v : bound[q]) o->bound_idx.push_back(v + index_base);
  }
  *out = o.release();
  return HS_OK;
  HS_TRY_END
}

extern "C" int32_t hs_nd_elimtree(const hs_ordering* o, hs_elimtree* et) {
  if (!o || !et) return hs_fail(HS_EARG, "hs_nd_elimtree: null argument");
  et->nnodes = o->nnodes;
  et->fathers = o->fathers.data(); et->lsons = o->lsons.data(); et->rsons = o->rsons.data();
  et->inter_ptr = o->inter_ptr.data(); et->inter_idx = o->inter_idx.data();
  et->bound_ptr = o->bound_ptr.data(); et->bound_idx = o->bound_idx.data();
  et->index_base = o->index_base;
  return HS_OK;
}

extern "C" int32_t hs_nd_free(hs_ordering* o) {
  delete o;
  return HS_OK;
}

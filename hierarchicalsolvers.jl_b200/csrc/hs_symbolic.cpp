// Symbolic phase on the host: parse_elimtree + symfact! + postorder + permuted! in one pass.
//
// Takes over reference src/nesteddissection.jl:105-148 (parse_elimtree), :29-69 (symfact!/_symfact!),
// :73-79 (postorder) and :82-88 (permuted!), in the order the driver calls them (test/rungmres.jl:15-19).
// The reference finds child→parent positions with `findall(in(nd.int), nd.left.bnd)`, an O(|parent|·|child|)
// scan per node; here membership is a stamped lookup table over the DOFs, so the whole phase is linear in the
// total index volume.
#include "hs_internal.h"

#include <algorithm>
#include <cstring>

struct hs_symbolic {
  int64_t nnodes = 0, n = 0, depth = 0;
  int32_t index_base = 0;
  std::vector<int64_t> left, right;
  std::vector<int64_t> int_ptr, int_idx, bnd_ptr, bnd_idx, iloc_ptr, iloc_idx, bloc_ptr, bloc_idx;
  std::vector<int64_t> perm;     // perm[new] = old  (index_base applied)
  std::vector<int64_t> orig_id;  // post-order position -> node id in the caller's numbering (0-based)
};

extern "C" int32_t hs_symfact(const hs_elimtree* et, int32_t apply_postorder, hs_symbolic** out) {
  HS_TRY_BEGIN
  if (!et || !out) return hs_fail(HS_EARG, "hs_symfact: null argument");
  const int64_t nn = et->nnodes;
  const int64_t base = et->index_base;
  if (nn <= 0) return hs_fail(HS_EARG, "hs_symfact: empty tree");
  // exactly one root (nesteddissection.jl:110-111)
  int64_t root = -1, nroots = 0;
  for (int64_t i = 0; i < nn; ++i)
    if (et->fathers[i] == -1) { root = i; ++nroots; }
  if (nroots != 1) return hs_fail(HS_EARG, "found either less than or more than one root.");
  auto child = [&](const int64_t* sons, int64_t i) -> int64_t {
    int64_t c = sons[i];
    if (c == -1) return -1;
    c -= base;
    if (c < 0 || c >= nn) throw hs_error(HS_EARG, "hs_symfact: child id out of range");
    return c;
  };
  // iterative post-order from the root (parse_elimtree walks the same way with two stacks, :114-147)
  std::vector<int64_t> order;
  order.reserve(nn);
  {
    std::vector<std::pair<int64_t, int>> st;
    st.push_back({root, 0});
    std::vector<char> seen(nn, 0);
    while (!st.empty()) {
      auto& top = st.back();
      int64_t i = top.first;
      if (top.second == 0) {
        if (seen[i]) throw hs_error(HS_EARG, "hs_symfact: elimination tree has a cycle or shared node");
        seen[i] = 1;
        top.second = 1;
        int64_t l = child(et->lsons, i);
        if (l >= 0) { st.push_back({l, 0}); continue; }
      }
      if (top.second == 1) {
        top.second = 2;
        int64_t r = child(et->rsons, i);
        if (r >= 0) { st.push_back({r, 0}); continue; }
      }
      order.push_back(i);
      st.pop_back();
    }
  }
  const int64_t m = (int64_t)order.size();
  std::vector<int64_t> newid(nn, -1);
  for (int64_t k = 0; k < m; ++k) newid[order[k]] = k;

  auto* S = new hs_symbolic();
  std::unique_ptr<hs_symbolic> guard(S);
  S->nnodes = m;
  S->index_base = (int32_t)base;
  S->orig_id = order;
  S->left.assign(m, -1);
  S->right.assign(m, -1);
  S->int_ptr.assign(m + 1, 0);
  S->bnd_ptr.assign(m + 1, 0);
  S->iloc_ptr.assign(m + 1, 0);
  S->bloc_ptr.assign(m + 1, 0);

  // largest DOF id, for the stamp table
  int64_t maxdof = -1;
  for (int64_t k = 0; k < m; ++k) {
    int64_t i = order[k];
    for (int64_t p = et->inter_ptr[i]; p < et->inter_ptr[i + 1]; ++p) maxdof = std::max(maxdof, et->inter_idx[p] - base);
    for (int64_t p = et->bound_ptr[i]; p < et->bound_ptr[i + 1]; ++p) maxdof = std::max(maxdof, et->bound_idx[p] - base);
  }
  std::vector<int64_t> stamp(maxdof + 1, -1);
  std::vector<char> kind(maxdof + 1, 0);

  // per-node lists in post-order; children are complete before their parent is visited
  std::vector<std::vector<int64_t>> nint(m), nbnd(m), iloc(m), bloc(m);
  std::vector<int64_t> height(m, 1);
  for (int64_t k = 0; k < m; ++k) {
    const int64_t i = order[k];
    const int64_t l = child(et->lsons, i), r = child(et->rsons, i);
    const int64_t i0 = et->inter_ptr[i], i1 = et->inter_ptr[i + 1];
    const int64_t b0 = et->bound_ptr[i], b1 = et->bound_ptr[i + 1];
    for (int64_t p = i0; p < i1; ++p)
      if (et->inter_idx[p] - base < 0) throw hs_error(HS_EARG, "hs_symfact: DOF id below index_base");
    for (int64_t p = b0; p < b1; ++p)
      if (et->bound_idx[p] - base < 0) throw hs_error(HS_EARG, "hs_symfact: DOF id below index_base");
    if (l < 0 && r < 0) {  // leaf: sets are taken as given (:117-118)
      nint[k].assign(et->inter_idx + i0, et->inter_idx + i1);
      nbnd[k].assign(et->bound_idx + b0, et->bound_idx + b1);
      for (auto& v : nint[k]) v -= base;
      for (auto& v : nbnd[k]) v -= base;
      continue;
    }
    for (int64_t p = i0; p < i1; ++p) { stamp[et->inter_idx[p] - base] = k; kind[et->inter_idx[p] - base] = 1; }
    for (int64_t p = b0; p < b1; ++p) { stamp[et->bound_idx[p] - base] = k; kind[et->bound_idx[p] - base] = 2; }
    std::vector<int64_t> bndr_tmp, intr_tmp;
    for (int side = 0; side < 2; ++side) {
      const int64_t c = side == 0 ? l : r;
      if (c < 0) continue;
      const int64_t ck = newid[c];
      (side == 0 ? S->left[k] : S->right[k]) = ck;
      height[k] = std::max(height[k], height[ck] + 1);
      const auto& cb = nbnd[ck];
      for (int64_t t = 0; t < (int64_t)cb.size(); ++t) {  // findall(in(nd.int), child.bnd) / findall(in(nd.bnd), child.bnd) :42-43,:54-55
        const int64_t dof = cb[t];
        if (dof <= maxdof && stamp[dof] == k) {
          if (kind[dof] == 1) { iloc[ck].push_back(t); nint[k].push_back(dof); }
          else { bloc[ck].push_back(t); nbnd[k].push_back(dof); }
        }
      }
    }
    // nd.int = [intl; intr], nd.bnd = [bndl; bndr] (:64-65) — the loop above appended left first, then right
  }
  // root: nd_loc.int = 1:length(nd.bnd), nd_loc.bnd = [] (:31-32)
  iloc[m - 1].resize(nbnd[m - 1].size());
  for (size_t t = 0; t < iloc[m - 1].size(); ++t) iloc[m - 1][t] = (int64_t)t;
  bloc[m - 1].clear();
  S->depth = height[m - 1];

  // postorder(nd) (:73-79): every node's int in post-order, then the root's bnd
  for (int64_t k = 0; k < m; ++k) S->perm.insert(S->perm.end(), nint[k].begin(), nint[k].end());
  S->perm.insert(S->perm.end(), nbnd[m - 1].begin(), nbnd[m - 1].end());
  S->n = (int64_t)S->perm.size();
  if (apply_postorder) {
    // permuted!(nd, invperm(perm)) (:82-88, rungmres.jl:19)
    std::vector<int64_t> iperm(maxdof + 1, -1);
    for (int64_t t = 0; t < S->n; ++t) {
      const int64_t old = S->perm[t];
      if (iperm[old] != -1) return hs_fail(HS_EARG, "hs_symfact: post-order is not a permutation (a DOF is eliminated twice)");
      iperm[old] = t;
    }
    for (int64_t k = 0; k < m; ++k) {
      for (auto& v : nint[k]) v = iperm[v];
      for (auto& v : nbnd[k]) {
        if (iperm[v] < 0) return hs_fail(HS_EARG, "hs_symfact: boundary DOF never eliminated");
        v = iperm[v];
      }
    }
  }
  auto flatten = [&](std::vector<std::vector<int64_t>>& src, std::vector<int64_t>& ptr, std::vector<int64_t>& idx) {
    int64_t tot = 0;
    for (int64_t k = 0; k < m; ++k) { ptr[k] = tot; tot += (int64_t)src[k].size(); }
    ptr[m] = tot;
    idx.resize(tot);
    for (int64_t k = 0; k < m; ++k) {
      int64_t* d = idx.data() + ptr[k];
      for (size_t t = 0; t < src[k].size(); ++t) d[t] = src[k][t] + base;
    }
  };
  flatten(nint, S->int_ptr, S->int_idx);
  flatten(nbnd, S->bnd_ptr, S->bnd_idx);
  flatten(iloc, S->iloc_ptr, S->iloc_idx);
  flatten(bloc, S->bloc_ptr, S->bloc_idx);
  for (auto& v : S->perm) v += base;
  *out = guard.release();
  return HS_OK;
  HS_TRY_END
}

extern "C" int32_t hs_symbolic_tree(const hs_symbolic* s, hs_tree* t) {
  if (!s || !t) return hs_fail(HS_EARG, "hs_symbolic_tree: null argument");
  t->nnodes = s->nnodes;
  t->left = s->left.data();
  t->right = s->right.data();
  t->int_ptr = s->int_ptr.data();
  t->int_idx = s->int_idx.data();
  t->bnd_ptr = s->bnd_ptr.data();
  t->bnd_idx = s->bnd_idx.data();
  t->iloc_ptr = s->iloc_ptr.data();
  t->iloc_idx = s->iloc_idx.data();
  t->bloc_ptr = s->bloc_ptr.data();
  t->bloc_idx = s->bloc_idx.data();
  t->index_base = s->index_base;
  return HS_OK;
}

extern "C" int32_t hs_symbolic_perm(const hs_symbolic* s, const int64_t** perm, int64_t* n) {
  if (!s || !perm || !n) return hs_fail(HS_EARG, "hs_symbolic_perm: null argument");
  *perm = s->perm.data();
  *n = s->n;
  return HS_OK;
}

extern "C" int32_t hs_symbolic_depth(const hs_symbolic* s, int64_t* depth) {
  if (!s || !depth) return hs_fail(HS_EARG, "hs_symbolic_depth: null argument");
  *depth = s->depth;
  return HS_OK;
}

extern "C" int32_t hs_symbolic_free(hs_symbolic* s) {
  delete s;
  return HS_OK;
}

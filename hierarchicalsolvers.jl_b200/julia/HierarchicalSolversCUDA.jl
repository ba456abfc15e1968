# HierarchicalSolversCUDA.jl — the Julia-side binding of libhsolve_cuda for bonevbs/HierarchicalSolvers.jl.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The same C ABI is exercised from Python
# (hierarchicalsolvers.jl_b200/_lib.py, tests/); this file is what a maintainer of the reference adds to route
# `factor` / `ldiv!` to the GPU.  It keeps the reference's host-side symbolic phase (nesteddissection.jl) and replaces
# storage + numerics (blockmatrix.jl, factornode.jl, factorization.jl) by device-resident fronts.
#
#   using HierarchicalSolvers, HierarchicalSolversCUDA
#   A, b, nd = read_problem(path); nd, nd_loc = symfact!(nd)
#   perm = postorder(nd); A = permute(A, perm, perm); nd = permuted!(nd, invperm(perm))
#   F = cufactor(A, nd, nd_loc; swlevel = 0)            # instead of factor(...)
#   x, ch = gmres(A, b; Pr = F, reltol = 1e-9, restart = 30, log = true, maxiter = 30)
#
# To run the reference's own driver unchanged (test/rungmres.jl calls `factor(A, nd, nd_loc; ...)`), call
# `HierarchicalSolversCUDA.activate!()` once after `using`: it adds a more specific `factor` method for
# SparseMatrixCSC{Float64/ComplexF64, Int} that routes to `cufactor` (`deactivate!()` removes it again).
module HierarchicalSolversCUDA

using LinearAlgebra, SparseArrays
using HierarchicalSolvers
import LinearAlgebra: ldiv!
import HierarchicalSolvers: maxrank, isleaf, isbranch

export CuFactorNode, cufactor, noderanks, nested_dissection, hssrank, activate!, deactivate!

const libhsolve = get(ENV, "LIBHSOLVE_CUDA", "libhsolve_cuda")

# ---- status codes → the exceptions the reference raises (include/hsolve_cuda.h) -------------------------------
const HS_OK, HS_EARG, HS_EDIM, HS_ETREE, HS_ESINGULAR, HS_ECUDA, HS_ENOMEM, HS_ENOTIMPL, HS_ESIZE = 0:8
function check(rc::Int32)
  rc == HS_OK && return
  msg = unsafe_string(ccall((:hs_last_error, libhsolve), Cstring, ()))
  rc == HS_EARG      && throw(ArgumentError(msg))
  rc == HS_EDIM      && throw(DimensionMismatch(msg))
  rc == HS_ESINGULAR && throw(SingularException(0))
  rc == HS_ENOMEM    && throw(OutOfMemoryError())
  throw(ErrorException(msg))     # HS_ETREE is the reference's ErrorException (factorization.jl:25)
end

# ---- plain-old-data mirrors of the ABI structs ----------------------------------------------------------------
struct HsOpts             # SolverOptions, HierarchicalSolvers.jl:30-40, followed by the library's extensions (hsolve_cuda.h)
  swlevel::Int64; swsize::Int64; atol::Float64; rtol::Float64; c_tol::Float64
  leafsize::Int64; kest::Int64; stepsize::Int64; verbose::Int32; subtree::Int32
  hss::Int32; pad0::Int32                         # 1: Schur complements of compressed nodes stored as HSS (the reference's behaviour)
  sketch_omega::Ptr{Cvoid}; sketch_psi::Ptr{Cvoid} # host-supplied Gaussian test matrices (C_NULL: drawn on the device)
  sketch_rows::Int64; sketch_cols::Int64; sketch_seed::UInt64
end
# `sketches = (Ω, Ψ)`: e.g. `Random.seed!(123); Ω = randn(T, nbmax, kmax); Ψ = randn(T, nbmax, kmax)` reproduces a run of the
# reference that draws the same matrices (test/rungmres.jl:7); the caller keeps them alive during the call (GC.@preserve)
HsOpts(o::SolverOptions; hss::Bool = true, sketches = nothing, seed::Integer = 123) =
  HsOpts(o.swlevel, o.swsize, o.atol, o.rtol, o.c_tol, o.leafsize, o.kest, o.stepsize, o.verbose, 0, hss, 0,
         sketches === nothing ? C_NULL : pointer(sketches[1]), sketches === nothing ? C_NULL : pointer(sketches[2]),
         sketches === nothing ? 0 : size(sketches[1], 1), sketches === nothing ? 0 : size(sketches[1], 2), UInt64(seed))

struct HsTree
  nnodes::Int64
  left::Ptr{Int64}; right::Ptr{Int64}
  int_ptr::Ptr{Int64}; int_idx::Ptr{Int64}; bnd_ptr::Ptr{Int64}; bnd_idx::Ptr{Int64}
  iloc_ptr::Ptr{Int64}; iloc_idx::Ptr{Int64}; bloc_ptr::Ptr{Int64}; bloc_idx::Ptr{Int64}
  index_base::Int32
end

# flatten the two parallel BinaryNode trees (nd, nd_loc) of symfact! into post-order ragged arrays
function flatten(nd::NestedDissection, nd_loc::NestedDissection)
  left = Int64[]; right = Int64[]
  ip = Int64[0]; ii = Int64[]; bp = Int64[0]; bi = Int64[]
  lip = Int64[0]; lii = Int64[]; lbp = Int64[0]; lbi = Int64[]
  function rec(x, xl)
    l = isnothing(x.left) ? -1 : rec(x.left, xl.left)
    r = isnothing(x.right) ? -1 : rec(x.right, xl.right)
    push!(left, l); push!(right, r)
    append!(ii, x.int); push!(ip, length(ii)); append!(bi, x.bnd); push!(bp, length(bi))
    append!(lii, xl.int); push!(lip, length(lii)); append!(lbi, xl.bnd); push!(lbp, length(lbi))
    return length(left)             # 1-based node id in post-order
  end
  rec(nd, nd_loc)
  return (; left, right, ip, ii, bp, bi, lip, lii, lbp, lbi)
end

const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function context()
  if CTX[] == C_NULL
    check(ccall((:hs_create, libhsolve), Int32, (Ref{Ptr{Cvoid}}, Int32), CTX, 0))
  end
  CTX[]
end

# ---- the factorization object -------------------------------------------------------------------------------
mutable struct CuFactorNode{T} <: Factorization{T}
  handle::Ptr{Cvoid}
  n::Int
  node::Int                  # post-order id (0-based) of this node, root = nnodes-1
  nd::NestedDissection
  nd_loc::NestedDissection
  root::Union{CuFactorNode{T}, Nothing}   # the node that owns the handle (nothing for the root itself)
  lefts::Vector{Int64}                     # post-order child ids (1-based, -1 = none) from `flatten`
  rights::Vector{Int64}
end
Base.eltype(::CuFactorNode{T}) where T = T
Base.size(F::CuFactorNode) = (getfield(F, :n), getfield(F, :n))
Base.show(io::IO, F::CuFactorNode) = print(io, "CuFactorNode{$(eltype(F))}")

dtype_code(::Type{Float64}) = Int32(0)
dtype_code(::Type{ComplexF64}) = Int32(1)

"""
    cufactor(A, nd, nd_loc, opts = SolverOptions(); kw...) -> CuFactorNode

Drop-in for `factor` (src/factorization.jl:5-11).  Copies `A` and the tree to the GPU and factors there.
"""
function cufactor(A::SparseMatrixCSC{T,Int}, nd::NestedDissection, nd_loc::NestedDissection,
                  opts::SolverOptions = SolverOptions(); hss::Bool = true, sketches = nothing, seed::Integer = 123,
                  args...) where T <: Union{Float64, ComplexF64}
  opts = copy(opts; args...)
  HierarchicalSolvers.chkopts!(opts)
  t = flatten(nd, nd_loc)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  sketches === nothing || (sketches = (Matrix{T}(sketches[1]), Matrix{T}(sketches[2])))
  GC.@preserve A t sketches begin
    tree = HsTree(length(t.left), pointer(t.left), pointer(t.right), pointer(t.ip), pointer(t.ii), pointer(t.bp), pointer(t.bi),
                  pointer(t.lip), pointer(t.lii), pointer(t.lbp), pointer(t.lbi), Int32(1))
    rc = ccall((:hs_factor, libhsolve), Int32,
               (Ptr{Cvoid}, Int32, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Ref{HsTree}, Ref{HsOpts}, Int32, Ref{Ptr{Cvoid}}),
               context(), dtype_code(T), size(A, 1), A.colptr, A.rowval, A.nzval, tree,
               HsOpts(opts; hss = hss, sketches = sketches, seed = seed), Int32(0), h)
  end
  if rc != HS_OK
    h[] != C_NULL && ccall((:hs_factor_free, libhsolve), Int32, (Ptr{Cvoid},), h[])
    check(rc)
  end
  F = CuFactorNode{T}(h[], size(A, 1), length(t.left) - 1, nd, nd_loc, nothing, t.left, t.right)
  finalizer(F -> ccall((:hs_factor_free, libhsolve), Int32, (Ptr{Cvoid},), getfield(F, :handle)), F)
  return F
end

# `factor(A, nd, nd_loc, opts; kw...)` itself (src/factorization.jl:5): a method on the concrete matrix types this library
# handles is more specific than the reference's `SparseMatrixCSC{T}` method, so after activate!() the reference's driver
# (test/rungmres.jl:32,39) reaches the GPU without an edit.  Defined at run time (extending another package's function on
# its own types is not allowed while this module precompiles).
function activate!()
  @eval HierarchicalSolvers.factor(A::SparseMatrixCSC{T,Int}, nd::HierarchicalSolvers.NestedDissection,
                                   nd_loc::HierarchicalSolvers.NestedDissection,
                                   opts::HierarchicalSolvers.SolverOptions = HierarchicalSolvers.SolverOptions();
                                   args...) where {T <: Union{Float64, ComplexF64}} = cufactor(A, nd, nd_loc, opts; args...)
  nothing
end
function deactivate!()
  for T in (Float64, ComplexF64)
    m = which(HierarchicalSolvers.factor, Tuple{SparseMatrixCSC{T,Int}, HierarchicalSolvers.NestedDissection, HierarchicalSolvers.NestedDissection})
    m.module === @__MODULE__() && Base.delete_method(m)
  end
  nothing
end

# ---- ldiv!  (src/factornode.jl:62-74) -----------------------------------------------------------------------
# 3-argument forms write into C.  The 2-argument form is genuinely in place here (the reference's allocates and
# leaves B untouched, factornode.jl:62 — see SURVEY F6); IterativeSolvers' right-preconditioned update relies on it.
function ldiv!(C::StridedVecOrMat{T}, F::CuFactorNode{T}, B::StridedVecOrMat{T}) where T
  n = getfield(F, :n)
  getfield(F, :root) === nothing || throw(ArgumentError("ldiv! is defined on the root of the factor tree"))
  size(B, 1) == n || throw(DimensionMismatch("B has $(size(B,1)) rows, expected $n"))
  Bc = Matrix{T}(reshape(B, n, :))            # contiguous staging copy (gmres hands in SubArray columns)
  GC.@preserve Bc check(ccall((:hs_solve, libhsolve), Int32, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int32),
                              getfield(F, :handle), size(Bc, 2), Bc, n, Bc, n, Int32(0)))
  copyto!(C, reshape(Bc, size(C)))
  return C
end
ldiv!(F::CuFactorNode{T}, B::StridedVecOrMat{T}) where T = ldiv!(B, F, B)
Base.:\(F::CuFactorNode{T}, B::StridedVecOrMat{T}) where T = ldiv!(similar(B), F, B)

function maxrank(F::CuFactorNode)             # src/factornode.jl:49-57
  r = Ref{Int64}(0)
  check(ccall((:hs_maxrank, libhsolve), Int32, (Ptr{Cvoid}, Ref{Int64}), getfield(F, :handle), r))
  return Int(r[])
end

# rank(F.L), rank(F.R) of a compressed node (LowRankMatrix, src/factorization.jl:173,179); (0, 0) when L, R are dense
function noderanks(F::CuFactorNode)
  rl = Ref{Int64}(0); rr = Ref{Int64}(0)
  check(ccall((:hs_node_rank, libhsolve), Int32, (Ptr{Cvoid}, Int64, Ref{Int64}, Ref{Int64}),
              getfield(F, :handle), getfield(F, :node), rl, rr))
  return Int(rl[]), Int(rr[])
end

# ---- elimination tree for a matrix that comes without one (the reference has no ordering code) ----------------
struct HsElimTree
  nnodes::Int64
  fathers::Ptr{Int64}; lsons::Ptr{Int64}; rsons::Ptr{Int64}
  inter_ptr::Ptr{Int64}; inter_idx::Ptr{Int64}; bound_ptr::Ptr{Int64}; bound_idx::Ptr{Int64}
  index_base::Int32
end

"""
    nested_dissection(A; nmax = 100) -> NestedDissection

Recursive METIS bisection of the pattern of `A + A'`; returns what `parse_elimtree` (src/nesteddissection.jl:105-148)
returns for the `elim_tree` of a problem file.
"""
function nested_dissection(A::SparseMatrixCSC{T,Int}; nmax::Int = 100) where T
  h = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve A check(ccall((:hs_nd_create, libhsolve), Int32,
                             (Int64, Ptr{Int64}, Ptr{Int64}, Int32, Int32, Int64, Ref{Ptr{Cvoid}}),
                             size(A, 1), A.colptr, A.rowval, Int32(0), Int32(1), nmax, h))
  et = Ref{HsElimTree}()
  check(ccall((:hs_nd_elimtree, libhsolve), Int32, (Ptr{Cvoid}, Ref{HsElimTree}), h[], et))
  e = et[]; nn = Int(e.nnodes)
  cp(p, m) = copy(unsafe_wrap(Array, p, m))
  fathers, lsons, rsons = cp(e.fathers, nn), cp(e.lsons, nn), cp(e.rsons, nn)
  ip, bp = cp(e.inter_ptr, nn + 1), cp(e.bound_ptr, nn + 1)
  ii, bi = cp(e.inter_idx, ip[end]), cp(e.bound_idx, bp[end])
  ccall((:hs_nd_free, libhsolve), Int32, (Ptr{Cvoid},), h[])
  ninter, nbound = diff(ip), diff(bp)
  inter = zeros(Int, max(maximum(ninter), 1), nn); bound = zeros(Int, max(maximum(nbound), 1), nn)
  for k in 1:nn
    inter[1:ninter[k], k] = ii[ip[k]+1:ip[k+1]]
    bound[1:nbound[k], k] = bi[bp[k]+1:bp[k+1]]
  end
  return HierarchicalSolvers.parse_elimtree(fathers, lsons, rsons, ninter, inter, nbound, bound)
end

# ---- FactorNode fields, copied from the device on access (src/factornode.jl:8-22) ----------------------------
const WHICH = Dict(:D => 0, :S => 1, :L => 2, :R => 3)
function Base.getproperty(F::CuFactorNode{T}, s::Symbol) where T
  if haskey(WHICH, s)
    dims = zeros(Int64, 2)
    check(ccall((:hs_node_get, libhsolve), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{Cvoid}, Ptr{Int64}),
                getfield(F, :handle), getfield(F, :node), Int32(WHICH[s]), C_NULL, dims))
    M = Matrix{T}(undef, dims[1], dims[2])
    check(ccall((:hs_node_get, libhsolve), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{Cvoid}, Ptr{Int64}),
                getfield(F, :handle), getfield(F, :node), Int32(WHICH[s]), M, dims))
    return M
  elseif s === :int;      return getfield(F, :nd).int
  elseif s === :bnd;      return getfield(F, :nd).bnd
  elseif s === :int_loc;  return getfield(F, :nd_loc).int
  elseif s === :bnd_loc;  return getfield(F, :nd_loc).bnd
  elseif s === :left || s === :right
    # children (src/factornode.jl:21-22) as lazy views on the same device factorization: same handle, the child's node id
    # and its sub-trees of (nd, nd_loc); the root is kept alive through `root` so the finalizer cannot run under a view
    c = (s === :left ? getfield(F, :lefts) : getfield(F, :rights))[getfield(F, :node) + 1]
    c < 0 && return nothing
    nd, ndl = getfield(F, :nd), getfield(F, :nd_loc)
    owner = something(getfield(F, :root), F)
    return CuFactorNode{T}(getfield(F, :handle), getfield(F, :n), Int(c) - 1, s === :left ? nd.left : nd.right,
                           s === :left ? ndl.left : ndl.right, owner, getfield(F, :lefts), getfield(F, :rights))
  else
    return getfield(F, s)
  end
end

# hssrank(F.S) of a compressed node (src/factornode.jl:53); 0 when S is dense
function hssrank(F::CuFactorNode)
  n = Ref{Int64}(0)
  check(ccall((:hs_hss_info, libhsolve), Int32, (Ptr{Cvoid}, Int64, Ref{Int64}, Ptr{Int64}), getfield(F, :handle), getfield(F, :node), n, C_NULL))
  n[] == 0 && return 0
  info = zeros(Int64, 8, n[])
  check(ccall((:hs_hss_info, libhsolve), Int32, (Ptr{Cvoid}, Int64, Ref{Int64}, Ptr{Int64}), getfield(F, :handle), getfield(F, :node), n, info))
  r = 0
  for t in 1:n[]
    info[8, t] == 1 && continue                       # leaf
    l, rgt = info[3, t] + 1, info[4, t] + 1
    r = max(r, info[5, l], info[6, l], info[5, rgt], info[6, rgt])   # dimensions of B12 / B21
  end
  return Int(r)
end
isleaf(F::CuFactorNode) = HierarchicalSolvers.isleaf(getfield(F, :nd))
isbranch(F::CuFactorNode) = HierarchicalSolvers.isbranch(getfield(F, :nd))

end # module

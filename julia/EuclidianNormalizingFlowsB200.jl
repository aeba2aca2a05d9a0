# EuclidianNormalizingFlowsB200.jl -- the Julia-side shim a maintainer of
# EuclidianNormalizingFlows.jl adds to route device-resident sample matrices
# through libenf_b200.so.  It adds *more specific methods* to the package's own
# generic functions for a new matrix type; nothing in the package is modified and
# host `Matrix` inputs keep using the reference's CPU methods.
#
# NOT RUNNABLE IN THE BUILD IMAGE (no Julia there): kept small and mechanical.
# Every ccall matches a declaration in include/enf_b200.h; tests/ bind the same
# symbols through ctypes (euclidiannormalizingflows.jl_b200/_lib.py).
module EuclidianNormalizingFlowsB200

using EuclidianNormalizingFlows
using EuclidianNormalizingFlows: CenterStretch, CenterContract, JohnsonTrafo, JohnsonTrafoInv,
    ScaleShiftTrafo, HouseholderTrafo
import EuclidianNormalizingFlows: mvnormal_negll_trafo, mvnormal_negll_trafograd
import ChangesOfVariables: with_logabsdet_jacobian
using LinearAlgebra: Adjoint

const libenf = get(ENV, "ENF_B200_LIB", "libenf_b200.so")
const ENF_NEGLL_ZYGOTE_PRIMAL = Cint(1)

struct EnfError <: Exception
    code::Cint
    msg::String
end
check(rc::Cint, ctx = C_NULL) = rc == 0 ? nothing :
    throw(EnfError(rc, unsafe_string(ccall((:enf_last_error, libenf), Cstring, (Ptr{Cvoid},), ctx))))

# ---- context (one per Julia process / GPU) ---------------------------------------
mutable struct Context
    handle::Ptr{Cvoid}
    chains::Dict{Any,Ptr{Cvoid}}
    function Context(device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:enf_init, libenf), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
        ctx = new(h[], Dict{Any,Ptr{Cvoid}}())
        finalizer(ctx) do c
            foreach(ch -> ccall((:enf_chain_destroy, libenf), Cint, (Ptr{Cvoid},), ch), values(c.chains))
            ccall((:enf_destroy, libenf), Cint, (Ptr{Cvoid},), c.handle)
        end
    end
end
const default_ctx = Ref{Union{Nothing,Context}}(nothing)
context() = something(default_ctx[], (default_ctx[] = Context(0)))

# ---- device matrix: D x N, column-major, ld = D (the memory of a Matrix{T}) -------
mutable struct B200Matrix{T<:Union{Float32,Float64}} <: DenseMatrix{T}
    ctx::Context
    ptr::Ptr{Cvoid}
    dims::Tuple{Int,Int}
    owner::Any                       # parent matrix for column views, else nothing
end
Base.size(x::B200Matrix) = x.dims
function B200Matrix{T}(ctx::Context, D::Integer, N::Integer) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:enf_alloc, libenf), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx.handle, D * N * sizeof(T), p), ctx.handle)
    x = B200Matrix{T}(ctx, p[], (D, N), nothing)
    finalizer(m -> ccall((:enf_free, libenf), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), m.ctx.handle, m.ptr), x)
end
function B200Matrix(X::Matrix{T}, ctx::Context = context()) where {T}
    x = B200Matrix{T}(ctx, size(X)...)
    GC.@preserve X check(ccall((:enf_h2d, libenf), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{T}, Csize_t),
                               ctx.handle, x.ptr, X, sizeof(X)), ctx.handle)
    check(ccall((:enf_sync, libenf), Cint, (Ptr{Cvoid},), ctx.handle), ctx.handle)
    x
end
function Base.Array(x::B200Matrix{T}) where {T}
    X = Matrix{T}(undef, x.dims...)
    GC.@preserve X check(ccall((:enf_d2h, libenf), Cint, (Ptr{Cvoid}, Ptr{T}, Ptr{Cvoid}, Csize_t),
                               x.ctx.handle, X, x.ptr, sizeof(X)), x.ctx.handle)
    X
end
# flatview(batch) of a partitioned nestedview(X) is a contiguous column range
# (src/optimize_whitening.jl:32,38): a view is pointer + offset, no copy.
Base.view(x::B200Matrix{T}, ::Colon, r::UnitRange{Int}) where {T} =
    B200Matrix{T}(x.ctx, x.ptr + (first(r) - 1) * x.dims[1] * sizeof(T), (x.dims[1], length(r)), x)

# ---- trafo tree -> flat op list (innermost first), params expanded to length D ------
const Leaf = Union{CenterStretch,CenterContract,JohnsonTrafo,JohnsonTrafoInv,ScaleShiftTrafo,HouseholderTrafo}
flatten(f::Base.ComposedFunction) = vcat(flatten(f.inner), flatten(f.outer))
flatten(f::Leaf) = Any[f]
kind(::CenterStretch) = 0; kind(::CenterContract) = 1; kind(::JohnsonTrafo) = 2
kind(::JohnsonTrafoInv) = 3; kind(::ScaleShiftTrafo) = 4; kind(::HouseholderTrafo) = 5
nrefl(f::HouseholderTrafo) = size(f.V, 2); nrefl(::Leaf) = 0
expand(p::Real, D, T) = fill(T(p), D)
expand(p::AbstractVector, D, T) = (length(p) == D || throw(DimensionMismatch()); Vector{T}(p))
params(f::HouseholderTrafo, D, T) = Vector{T}(vec(f.V))
params(f::Leaf, D, T) = reduce(vcat, (expand(getfield(f, n), D, T) for n in fieldnames(typeof(f))))

struct EnfOp
    kind::Int32
    K::Int32
    params::Ptr{Cvoid}
end

function chain(ctx::Context, f, D::Int, ::Type{T}) where {T}
    leaves = flatten(f)
    packed = reduce(vcat, (params(l, D, T) for l in leaves))
    key = (T, D, Tuple((kind(l), nrefl(l)) for l in leaves))
    ch = get(ctx.chains, key, C_NULL)
    GC.@preserve packed begin
        if ch == C_NULL
            ops, off = EnfOp[], 0
            for l in leaves
                n = length(params(l, D, T))
                push!(ops, EnfOp(kind(l), nrefl(l), pointer(packed) + off * sizeof(T)))
                off += n
            end
            h = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:enf_chain_create, libenf), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{EnfOp}, Ref{Ptr{Cvoid}}),
                        ctx.handle, T === Float32 ? 0 : 1, D, length(ops), ops, h), ctx.handle)
            ch = ctx.chains[key] = h[]
        else
            check(ccall((:enf_chain_set_params, libenf), Cint, (Ptr{Cvoid}, Ptr{T}), ch, packed), ctx.handle)
        end
    end
    ch, leaves
end

# ---- the generic functions of the hot path -------------------------------------------
const Trafo = Union{Leaf,Base.ComposedFunction}

# The reference computes in float(promote_type(eltype(x), eltype(params)...)) (src/center_stretch.jl:5,
# src/johnson_trafo.jl:30, src/scale_shift_trafo.jl:15).  A chain is all-Float32 or all-Float64, so Float32 samples
# meeting Float64 parameters are widened on the device (enf_convert) - never the parameters narrowed.
paramtype(f::Leaf) = promote_type((eltype(getfield(f, n)) for n in fieldnames(typeof(f)))...)
paramtype(f::Base.ComposedFunction) = promote_type(paramtype(f.inner), paramtype(f.outer))
enf_dtype(::Type{Float32}) = Cint(0); enf_dtype(::Type{Float64}) = Cint(1)
function Base.convert(::Type{B200Matrix{P}}, x::B200Matrix{T}) where {P,T}
    P === T && return x
    y = B200Matrix{P}(x.ctx, size(x)...)
    check(ccall((:enf_convert, libenf), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64),
                x.ctx.handle, enf_dtype(P), y.ptr, enf_dtype(T), x.ptr, length(x)), x.ctx.handle)
    y
end
promoted(f::Trafo, x::B200Matrix{T}) where {T} = convert(B200Matrix{float(promote_type(T, paramtype(f)))}, x)

# (f::Trafo)(x): src/center_stretch.jl:37,61; johnson_trafo.jl:74,99; scale_shift_trafo.jl:15-16; householder_trafo.jl:156-157
apply(f::Trafo, x::B200Matrix) = _apply(f, promoted(f, x))
function _apply(f::Trafo, x::B200Matrix{T}) where {T}
    ch, _ = chain(x.ctx, f, size(x, 1), T)
    y = B200Matrix{T}(x.ctx, size(x)...)
    check(ccall((:enf_forward, libenf), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}), ch, x.ptr, size(x, 2), y.ptr), x.ctx.handle)
    y
end
for F in (:CenterStretch, :CenterContract, :JohnsonTrafo, :JohnsonTrafoInv, :ScaleShiftTrafo, :HouseholderTrafo)
    @eval (f::$F)(x::B200Matrix) = apply(f, x)
end
(f::Base.ComposedFunction)(x::B200Matrix) = apply(f, x)

# with_logabsdet_jacobian: src/center_stretch.jl:39,63; johnson_trafo.jl:76,101; scale_shift_trafo.jl:18;
# householder_trafo.jl:159-160.  ladj comes back as the 1 x N Adjoint row of src/abstract_trafo.jl:9.
with_logabsdet_jacobian(f::Trafo, x::B200Matrix) = _with_logabsdet_jacobian(f, promoted(f, x))
function _with_logabsdet_jacobian(f::Trafo, x::B200Matrix{T}) where {T}
    ch, _ = chain(x.ctx, f, size(x, 1), T)
    y = B200Matrix{T}(x.ctx, size(x)...)
    l = B200Matrix{T}(x.ctx, size(x, 2), 1)
    check(ccall((:enf_forward_ladj, libenf), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                ch, x.ptr, size(x, 2), y.ptr, l.ptr), x.ctx.handle)
    y, vec(Array(l))'
end

# src/optimize_whitening.jl:7-15
mvnormal_negll_trafo(f::Trafo, x::B200Matrix) = _mvnormal_negll_trafo(f, promoted(f, x))
function _mvnormal_negll_trafo(f::Trafo, x::B200Matrix{T}) where {T}
    ch, _ = chain(x.ctx, f, size(x, 1), T)
    out = Ref{Float64}(0)
    check(ccall((:enf_negll, libenf), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Float64}), ch, x.ptr, size(x, 2), out), x.ctx.handle)
    T(out[])
end

# packed gradient -> the nested NamedTuple Zygote returns, so Optimisers.update
# (src/optimize_whitening.jl:40) works unchanged
unpack(f::Base.ComposedFunction, g, D, pos) = (inner = unpack(f.inner, g, D, pos); outer = unpack(f.outer, g, D, pos); (outer = outer, inner = inner))
function unpack(f::HouseholderTrafo, g, D, pos)
    K = size(f.V, 2); v = reshape(g[pos[] .+ (1:D*K)], D, K); pos[] += D * K
    (V = v,)
end
function unpack(f::Leaf, g, D, pos)
    names = fieldnames(typeof(f))
    vals = map(names) do n
        v = g[pos[] .+ (1:D)]; pos[] += D
        getfield(f, n) isa Real ? sum(v) : v
    end
    NamedTuple{names}(vals)
end

# src/optimize_whitening.jl:18-22 (value as Zygote reports it: src/abstract_trafo.jl:30-33)
mvnormal_negll_trafograd(f::Trafo, x::B200Matrix) = _mvnormal_negll_trafograd(f, promoted(f, x))
function _mvnormal_negll_trafograd(f::Trafo, x::B200Matrix{T}) where {T}
    ch, leaves = chain(x.ctx, f, size(x, 1), T)
    np = Ref{Int64}(0)
    check(ccall((:enf_chain_num_params, libenf), Cint, (Ptr{Cvoid}, Ref{Int64}), ch, np), x.ctx.handle)
    g = Vector{T}(undef, np[])
    out = Ref{Float64}(0)
    GC.@preserve g check(ccall((:enf_negll_grad, libenf), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Ref{Float64}, Ptr{T}),
                               ch, x.ptr, size(x, 2), ENF_NEGLL_ZYGOTE_PRIMAL, out, g), x.ctx.handle)
    T(out[]), unpack(f, g, size(x, 1), Ref(0))
end

# The whole fit loop on the device (enf_optimize_whitening, SURVEY §8f n1): same arguments and result as
# optimize_whitening (src/optimize_whitening.jl:25-45) for an ADAGrad optimizer; two kernel launches per step.
function optimize_whitening_device(smpls::B200Matrix{T}, initial_trafo::Trafo; eta = 0.1f0, epsilon = eps(Float32),
                                   nbatches::Integer = 100, nepochs::Integer = 100) where {T}
    ch, _ = chain(smpls.ctx, initial_trafo, size(smpls, 1), T)
    np = Ref{Int64}(0)
    check(ccall((:enf_chain_num_params, libenf), Cint, (Ptr{Cvoid}, Ref{Int64}), ch, np), smpls.ctx.handle)
    nb = cld(size(smpls, 2), round(Int, size(smpls, 2) / nbatches))
    state = Vector{Float64}(undef, np[]); params = Vector{T}(undef, np[]); hist = Vector{Float64}(undef, nb * nepochs)
    nsteps = Ref{Int64}(0)
    GC.@preserve state params hist check(ccall((:enf_optimize_whitening, libenf), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Float64, Float64, Cint, Cint, Cint, Ptr{Float64}, Ptr{T}, Ptr{Float64}, Ref{Int64}),
        ch, smpls.ptr, size(smpls, 2), nbatches, nepochs, eta, epsilon, ENF_NEGLL_ZYGOTE_PRIMAL, 0, 1, state, params, hist, nsteps),
        smpls.ctx.handle)
    (params = params, optimizer_state = state, negll_history = hist[1:nsteps[]])   # packed like the C ABI; unpack() rebuilds the tree
end

# optimize_whitening itself (src/optimize_whitening.jl:25-45) needs one extra method so that
# `flatview(batch)` of a device matrix is a column view; everything else is the reference's loop:
#   smpls = B200Matrix(X); optimize_whitening(nestedview(smpls), initial_trafo, ADAGrad(); ...)
end # module

# ProtoGradB200.jl — Julia glue behind ProtoGrad's own extension seams.
#
# SOURCE ONLY: there is no Julia toolchain in the build image, so this file has never been executed.  What IS checked
# here (tests/test_julia_glue.py): every `ccall` names a symbol that include/protograd_b200.h declares, with the same
# number of arguments; PgOp / PgMlpDesc have the header's struct layouts; blocks and brackets balance.  The Python host mirror (protograd.jl_b200/) drives exactly the same C ABI and is what the GPU
# tests run; tests/abi_driver.py drives a whole step through pg_arena_* / pg_plan_* / pg_opt_* only, i.e. through the
# calls this file makes.
#
# It would live next to ext/ProtoGradZygoteExt.jl as a second package extension (Project.toml [extensions]) and adds,
# without touching or overwriting any method of the reference:
#   * DeviceModel{M} <: Model      a model whose leaves live in ONE device arena in vec(m) order (src/model.jl:33);
#                                  `to_device(m)` builds it, `to_host(dm)` brings it back
#   * eval_with_pullback(f, dm, ::Val{:B200})                       (src/gradient.jl:1-2 seam)
#   * Base.copyto!(dest::DeviceModel, bc::Broadcasted{Style{…}})    one kernel per dotted statement (src/model.jl:148-152)
#   * dot / zero / copy / similar / vec / length / getindex / setindex!   (src/model.jl:33,49-56,79-99,200-202)
#   * step!(state, grad) for the four optimizers                    (src/optimizers/*.jl)
#   * DataParallel: the flat-gradient exchange (pg_dp_*), for one Julia process per GPU
module ProtoGradB200

using ProtoGrad
using ProtoGrad: Model, allparams
using LinearAlgebra
using Statistics: mean
import NNlib
import Base.Broadcast: Broadcasted, Style, broadcasted

const LIB = joinpath(@__DIR__, "..", "protograd.jl_b200", "libprotograd_b200.so")

# ---- enums of include/protograd_b200.h ----------------------------------------------------------------------------
const PG_F32, PG_F64, PG_BF16, PG_F16, PG_I32 = Cint(0), Cint(1), Cint(2), Cint(3), Cint(4)
const PG_SCALAR_F64, PG_SCALAR2_F64, PG_MATH_FAST = Cint(1), Cint(2), Cint(4)
const PG_MATH_FP32, PG_MATH_BF16 = Cint(0), Cint(1)
const OP_INPUT, OP_PARAM, OP_LINEAR, OP_CONV2D, OP_RELU, OP_MAXPOOL, OP_MEANPOOL, OP_RESHAPE, OP_DROPOUT, OP_SOFTMAX, OP_LOSS_CE, OP_LOSS_MSE =
    ntuple(i -> Cint(i - 1), 12)
const BC_ARG, BC_CONST, BC_CONST_T, BC_CONST_F32 = Int32(1), Int32(2), Int32(3), Int32(4)
const BC_BIN = Dict{Any,Int32}(+ => 10, - => 11, * => 12, / => 13, max => 14, min => 15, ^ => 16)
const BC_UN = Dict{Any,Int32}(sqrt => 31, abs => 32, exp => 33, log => 34, abs2 => 35, NNlib.relu => 36, inv => 37)

code(::Type{Float32}) = PG_F32
code(::Type{Float64}) = PG_F64
code(::Type{Float16}) = PG_F16

function check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:pg_last_error, LIB), Cstring, ()))
    error("protograd_b200: $msg")            # the reference's only error idiom (src/gradient.jl:1)
end

mutable struct Context
    h::Ptr{Cvoid}
end
function Context(device::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pg_init, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
    ctx = Context(h[])
    finalizer(c -> ccall((:pg_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), ctx)
    return ctx
end
const CTX = Ref{Context}()
ctx() = (isassigned(CTX) || (CTX[] = Context(parse(Int, get(ENV, "LOCAL_RANK", "0")))); CTX[])
sync() = check(ccall((:pg_sync, LIB), Cint, (Ptr{Cvoid},), ctx().h))

# ---- arena + leaves -------------------------------------------------------------------------------------------------
mutable struct Arena{T}
    h::Ptr{Cvoid}               # pg_arena*
    sizes::Vector{Int64}
end
function Arena{T}(sizes::Vector{Int64}) where {T}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pg_arena_create, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ref{Ptr{Cvoid}}), ctx().h, code(T), length(sizes), sizes, h))
    a = Arena{T}(h[], sizes)
    finalizer(x -> ccall((:pg_arena_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), a)
    return a
end
"similar (mode 0) / zero (1) / copy (2) of an arena                      (src/model.jl:49-56)"
function like(a::Arena{T}, mode::Integer) where {T}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pg_arena_like, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), a.h, mode, h))
    b = Arena{T}(h[], a.sizes)
    finalizer(x -> ccall((:pg_arena_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), b)
    return b
end
function info(a::Arena)
    dt, nl, len, npad, base = Ref{Cint}(0), Ref{Cint}(0), Ref{Int64}(0), Ref{Int64}(0), Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pg_arena_info, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Cint}, Ref{Int64}, Ref{Int64}, Ref{Ptr{Cvoid}}), a.h, dt, nl, len, npad, base))
    return (length = len[], n_padded = npad[], base = base[])
end
base_ptr(a::Arena) = info(a).base
padded_length(a::Arena) = info(a).n_padded

"A Model leaf: a view at a fixed offset of the arena.  The struct definitions of src/layers.jl:15-18,38-41 stay as they
are — they are parametric in the array type."
struct DeviceLeaf{T,N} <: AbstractArray{T,N}
    arena::Arena{T}
    index::Int                  # 0-based leaf index in vec(m) order
    dims::NTuple{N,Int}
end
Base.size(l::DeviceLeaf) = l.dims
function ptr(l::DeviceLeaf)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pg_arena_leaf_ptr, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}, Ptr{Int64}), l.arena.h, l.index, p, C_NULL))
    return p[]
end
"host copy of one leaf (`grad.W` in test/test_models.jl:151)"
function Base.Array(l::DeviceLeaf{T,N}) where {T,N}
    out = Array{T,N}(undef, l.dims)
    GC.@preserve out check(ccall((:pg_arena_download, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), l.arena.h, l.index, pointer(out)))
    return out
end
# scalar host access is a device round trip (correctness-only path, src/model.jl:79-99)
Base.getindex(l::DeviceLeaf, i::Int) = Array(l)[i]
Base.getindex(l::DeviceLeaf{T,N}, I::Vararg{Int,N}) where {T,N} = Array(l)[I...]

"The device-backed model: `model` has the user's struct types with DeviceLeaf leaves; Function fields are shared."
struct DeviceModel{M<:Model,T} <: Model
    model::M
    arena::Arena{T}
end
inner(d::DeviceModel) = getfield(d, :model)
arena(d::DeviceModel) = getfield(d, :arena)
Base.getproperty(d::DeviceModel, s::Symbol) = (s === :model || s === :arena) ? getfield(d, s) : getproperty(inner(d), s)
(d::DeviceModel)(x) = inner(d)(x)
ProtoGrad._push_allparams!(c, ::Arena) = c                 # the arena handle is not a parameter (src/model.jl:9-24)
Base.length(d::DeviceModel) = Int(info(arena(d)).length)
Base.eltype(::DeviceModel{M,T}) where {M,T} = T

"Move a CPU model onto the device: one arena in vec(m) order, leaves become views."
function to_device(m::M) where {M<:Model}
    ps = allparams(m)
    T = eltype(first(ps))
    a = Arena{T}(Int64[length(p) for p in ps])
    flat = vec(m)                                               # src/model.jl:33
    GC.@preserve flat check(ccall((:pg_arena_upload, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), a.h, -1, pointer(flat)))
    return DeviceModel(rebuild(m, a), a)
end
# The structural map of src/model.jl:42-47 (`map_recursively`) rebuilds a struct as `T(fields...)` with T the CONCRETE
# parametric type of the source — `Linear{Matrix{Float32},Vector{Float32}}(leaf, leaf)` would `convert` the new leaves
# back to host Arrays — and maps a Tuple with `fun.(t)` (no recursion into the layers of a Compose).  Both are fine for
# similar / copy / zero, which keep the array types; changing the leaf type needs the type WRAPPER's constructor
# (`Linear(W, b)` infers the parameters) and its own recursion.  Field-less models (ReLU, SoftMax{D}, MaxPool{S},
# Dropout{P,D}) carry their configuration in type parameters only and are shared as they are.
remap(fun, f::Function) = f
remap(fun, a::AbstractArray) = fun(a)
remap(fun, t::Tuple) = map(x -> remap(fun, x), t)
function remap(fun, m::T) where {T<:Model}
    names = fieldnames(T)
    isempty(names) && return m
    return Base.typename(T).wrapper((remap(fun, getfield(m, k)) for k in names)...)
end
"same struct (wrapper) types, leaves replaced by views into `a`, numbered in allparams / vec(m) order (src/model.jl:9-24)"
function rebuild(m::Model, a::Arena{T}) where {T}
    k = Ref(-1)
    return remap(x -> (k[] += 1; DeviceLeaf{T,ndims(x)}(a, k[], size(x))), m)
end
"vec(m) of the device model: the whole arena in leaf order, padding skipped (src/model.jl:33)"
function Base.vec(d::DeviceModel{M,T}) where {M,T}
    flat = Vector{T}(undef, length(d))
    GC.@preserve flat check(ccall((:pg_arena_download, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), arena(d).h, -1, pointer(flat)))
    return flat
end
function to_host(d::DeviceModel{M,T}) where {M,T}
    flat = vec(d)
    off = Ref(0)
    return remap(x -> (n = length(x); r = reshape(flat[off[]+1:off[]+n], size(x)); off[] += n; r), inner(d))
end
Base.collect(d::DeviceModel) = vec(d)
Base.iterate(d::DeviceModel, st = (vec(d), 1)) = st[2] > length(st[1]) ? nothing : (st[1][st[2]], (st[1], st[2] + 1))

for (fun, mode) in ((:similar, 0), (:zero, 1), (:copy, 2))
    @eval function Base.$fun(d::DeviceModel)
        a = like(arena(d), $mode)
        return DeviceModel(rebuild(inner(d), a), a)
    end
end
function Base.getindex(d::DeviceModel{M,T}, i::Integer) where {M,T}
    r = Ref{Cdouble}(0)
    check(ccall((:pg_arena_get_scalar, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Cdouble}), arena(d).h, i - 1, r))
    return T(r[])
end
function Base.setindex!(d::DeviceModel, v, i::Integer)
    check(ccall((:pg_arena_set_scalar, LIB), Cint, (Ptr{Cvoid}, Int64, Cdouble), arena(d).h, i - 1, Float64(v)))
    return v
end
function LinearAlgebra.dot(a::DeviceModel{M,T}, b::DeviceModel{M,T}) where {M,T}
    r = Ref{Cdouble}(0)
    check(ccall((:pg_vec_dot, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Cdouble}),
                ctx().h, code(T), base_ptr(arena(a)), base_ptr(arena(b)), padded_length(arena(a)), r))
    return T(r[])                                           # the leaf eltype (test/test_models.jl:82)
end

# ---- model algebra: ONE kernel per dotted statement (src/model.jl:148-164) --------------------------------------------
# The nested Broadcasted tree of `dest .= f.(args...)` is compiled into the postfix program of pg_vec_expr: arena
# operands, scalars with Julia's promotion level (Float64 literal -> the op runs in Float64; Int or T -> stays in T),
# + - * / ^ max min sqrt abs exp log abs2 relu inv, and `x .^ 2` (literal_pow).  Any other function is an error: there
# is no CPU fallback.
struct Program{T}
    args::Vector{Ptr{Cvoid}}
    code::Vector{Int32}
    consts::Vector{Float64}
end
unref(x::Base.RefValue) = x[]
unref(x) = x
function emit!(p::Program{T}, d::DeviceModel) where {T}
    b = base_ptr(arena(d))
    k = findfirst(==(b), p.args)
    k === nothing && (push!(p.args, b); k = length(p.args))
    push!(p.code, BC_ARG | Int32((k - 1) << 8))
end
function emit!(p::Program{T}, c::Number) where {T}
    kind = c isa Float64 && T !== Float64 ? BC_CONST : (c isa Float32 && T === Float16 ? BC_CONST_F32 : BC_CONST_T)
    T === Float64 && (kind = BC_CONST)
    push!(p.consts, Float64(c))
    push!(p.code, kind | Int32((length(p.consts) - 1) << 8))
end
emit!(p::Program, r::Base.RefValue) = emit!(p, r[])
function emit!(p::Program, bc::Broadcasted)
    f, args = bc.f, bc.args
    if f === Base.literal_pow && unref(args[1]) === (^) && unref(args[3]) === Val(2)
        emit!(p, args[2]); push!(p.code, BC_UN[abs2]); return
    end
    if length(args) == 1 && f === (-)
        emit!(p, args[1]); push!(p.code, Int32(30)); return                  # PG_BC_NEG
    end
    if length(args) == 1 && haskey(BC_UN, f)
        emit!(p, args[1]); push!(p.code, BC_UN[f]); return
    end
    if length(args) >= 2 && haskey(BC_BIN, f)
        emit!(p, args[1])
        for a in args[2:end]                                                 # `a .+ b .+ c` parses as +(a, b, c)
            emit!(p, a); push!(p.code, BC_BIN[f])
        end
        return
    end
    error("protograd_b200: function $(f) is not available in a device broadcast (no CPU fallback)")
end
function Base.copyto!(dest::DeviceModel{M,T}, bc::Broadcasted{Style{DeviceModel{M,T}}}) where {M,T}
    p = Program{T}(Ptr{Cvoid}[], Int32[], Float64[])
    emit!(p, bc)
    a = arena(dest)
    check(ccall((:pg_vec_expr, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Int32}, Cint, Ptr{Float64}, Cint, Int64, Cint, Ptr{Int64}),
                ctx().h, code(T), base_ptr(a), p.args, length(p.args), p.code, length(p.code), p.consts, length(p.consts),
                padded_length(a), length(a.sizes), a.sizes))
    return dest
end
# `dest .= src` (nesterov_method.jl:34) and `dest .= c`
function Base.copyto!(dest::DeviceModel{M,T}, src::DeviceModel{M,T}) where {M,T}
    a = arena(dest)
    check(ccall((:pg_copy_d2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, base_ptr(a), base_ptr(arena(src)), padded_length(a) * sizeof(T)))
    return dest
end

# ---- optimizers: step! as one fused kernel (src/optimizers/*.jl) ------------------------------------------------------
isf64(s) = s isa Float64 ? PG_SCALAR_F64 : Cint(0)            # Julia promotion rule (SURVEY.md A.6)
bp(d::DeviceModel) = base_ptr(arena(d))
np(d::DeviceModel) = padded_length(arena(d))

# state.variable .-= state.stepsize .* gradient                     (gradient_descent.jl:15-17)
function ProtoGrad.step!(state::ProtoGrad.GradientDescentState{S,V}, g::V) where {S,T,M,V<:DeviceModel{M,T}}
    check(ccall((:pg_opt_gd, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Cint, Ptr{Cvoid}),
                ctx().h, code(T), bp(state.variable), bp(g), np(g), Float64(state.stepsize), isf64(state.stepsize), C_NULL))
    return state.variable
end
# adam.jl:32-41: `it` before the increment; m_hat / v_hat are state fields in the reference and are materialised
function ProtoGrad.step!(state::ProtoGrad.AdamState{R,V}, g::V) where {R,T,M,V<:DeviceModel{M,T}}
    o = state.optimizer
    check(ccall((:pg_opt_adam, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Cdouble, Cdouble, Cdouble, Int64, Cint, Ptr{Cvoid}),
                ctx().h, code(T), bp(state.variable), bp(state.m), bp(state.v), bp(state.m_hat), bp(state.v_hat),
                bp(g), np(g), Float64(state.stepsize), Float64(o.beta1), Float64(o.beta2), Float64(o.epsilon), state.it, isf64(state.stepsize), C_NULL))
    return state.it += 1                              # adam.jl:40
end
# d = mu*d - s*g ; p += d                                            (heavy_ball_method.jl:16-19)
function ProtoGrad.step!(state::ProtoGrad.HeavyBallMethodState{S,V}, g::V) where {S,T,M,V<:DeviceModel{M,T}}
    mu = state.optimizer.momentum
    flags = isf64(state.stepsize) | (mu isa Float64 ? PG_SCALAR2_F64 : Cint(0))
    check(ccall((:pg_opt_heavyball, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Cdouble, Cint, Ptr{Cvoid}),
                ctx().h, code(T), bp(state.variable), bp(state.step), bp(g), np(g), Float64(state.stepsize), Float64(mu), flags, C_NULL))
    return state.variable
end
# prev = p ; p -= s*g ; p += mu*(p - prev), mu popped from the stateful host-side iterator (nesterov_method.jl:33-39)
function ProtoGrad.step!(state::ProtoGrad.NesterovState{S,V,I}, g::V) where {S,T,M,V<:DeviceModel{M,T},I}
    mu = popfirst!(state.momentum_iterator)                                   # Float64
    check(ccall((:pg_opt_nesterov, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Cdouble, Cint, Ptr{Cvoid}),
                ctx().h, code(T), bp(state.variable), bp(state.variable_prev), bp(g), np(g), Float64(state.stepsize), Float64(mu), isf64(state.stepsize), C_NULL))
    return state.variable
end

# ---- tracing f(m): the ops the reference's layers call, recorded instead of executed ---------------------------------
# C layout == Julia memory: Julia x[in, B] is C X[B][in]; Julia w[kW,kH,Cin,Cout] is C w[Cout][Cin][kH][kW]: shapes are
# reversed when they cross the ABI.
mutable struct Node
    kind::Cint
    dtype::Cint
    ins::Vector{Int}            # 0-based node ids
    dims::Tuple                 # Julia dims
    grad_leaf::Int
    k::Int
    p::Float64
    denom::Float64
    data::Any                   # INPUT: host / device array; PARAM: DeviceLeaf
end
mutable struct Trace{T}
    nodes::Vector{Node}
    seen::IdDict{Any,Int}
    grad_arena::Ptr{Cvoid}      # arena handle of the differentiated model: its leaves get gradient leaf indices
end
const ACTIVE = Ref{Any}(nothing)
struct Tracer{T,N} <: AbstractArray{T,N}
    trace::Trace{T}
    id::Int
    dims::NTuple{N,Int}
    pending::Symbol             # :none, or :matmul / :conv waiting for its `.+ b` (layers.jl:31,59)
    px::Int                     # pending product: node ids of its input and of its weight leaf
    pw::Int
end
Tracer{T,N}(tr, id, dims, pending) where {T,N} = Tracer{T,N}(tr, id, dims, pending, -1, -1)
Base.size(t::Tracer) = t.dims
Base.getindex(::Tracer, I...) = error("a traced value has no elements on the host")
function push_node!(tr::Trace{T}, kind, ins, dims; dtype = code(T), grad_leaf = -1, k = 0, p = 0.0, denom = 0.0, data = nothing, pending = :none) where {T}
    push!(tr.nodes, Node(kind, dtype, collect(Int, ins), dims, grad_leaf, k, p, denom, data))
    return Tracer{T,length(dims)}(tr, length(tr.nodes) - 1, dims, pending)
end
function lift(tr::Trace{T}, x::DeviceLeaf) where {T}
    haskey(tr.seen, x) && return tr.seen[x]
    gl = x.arena.h == tr.grad_arena ? x.index : -1            # leaves of captured (frozen) models get no gradient (test_fine_tuning.jl:29-33)
    t = push_node!(tr, OP_PARAM, (), size(x); grad_leaf = gl, data = x)
    return tr.seen[x] = t.id
end
function lift(tr::Trace{T}, x::AbstractArray) where {T}
    haskey(tr.seen, x) && return tr.seen[x]
    dt = eltype(x) === Int32 ? PG_I32 : code(T)
    # the bytes handed to pg_upload must be a dense Array of T (or Int32 class indices): views, Bool one-hot matrices
    # (Flux.onehotbatch, examples/01_mnist_mlp.jl:17) and other eltypes are materialised once per step
    host = (x isa Array && (eltype(x) === T || eltype(x) === Int32)) ? x : Array{T}(x)
    t = push_node!(tr, OP_INPUT, (), size(x); dtype = dt, data = host)
    return tr.seen[x] = t.id
end
lift(::Trace, x::Tracer) = x.id
trace_of(a, b) = b isa Tracer ? b.trace : (a isa Tracer ? a.trace : ACTIVE[])        # `nothing` outside a trace

function traced_reshape(x::Tracer, dims::Tuple)
    x.pending === :none || error("protograd_b200: `W*x` / `conv(x, w)` must be followed by `.+ b` (layers.jl:31,59)")
    n = prod(x.dims)
    known = prod(d for d in dims if d isa Int; init = 1)
    nd = map(d -> d isa Colon ? n ÷ known : d, dims)
    prod(nd) == n || throw(DimensionMismatch("new dimensions $(nd) must be consistent with array size $n"))
    push_node!(x.trace, OP_RESHAPE, (x.id,), nd)
end
# one method per Base.reshape signature an AbstractArray would otherwise reach, so that none of them is ambiguous with
# Base's (`reshape(::AbstractArray, ::Int...)`, `(::AbstractArray, ::Dims)`, `(::AbstractArray, ::Union{Int,Colon}...)`,
# `(::AbstractArray, ::Tuple{Vararg{Union{Int,Colon}}})`): layers.jl:30,32 use the Colon forms, Reshape{S} the tuple form
Base.reshape(x::Tracer, dims::Int...) = traced_reshape(x, dims)
Base.reshape(x::Tracer, dims::Union{Int,Colon}...) = traced_reshape(x, dims)
Base.reshape(x::Tracer, dims::Dims) = traced_reshape(x, dims)
Base.reshape(x::Tracer, dims::Tuple{Vararg{Union{Int,Colon}}}) = traced_reshape(x, dims)
# l.W * x_reshaped  (layers.jl:31): recorded when the bias arrives
function Base.:*(W::DeviceLeaf{T,2}, x::AbstractMatrix{T}) where {T}
    tr = trace_of(W, x)
    tr === nothing && error("protograd_b200: a device model can only be called inside eval_with_pullback(f, m, :B200) or forward(m, x)")
    size(W, 2) == size(x, 1) || throw(DimensionMismatch("Linear expects $(size(W, 2)) features, got $(size(x, 1))"))
    return Tracer{T,2}(tr, -1, (size(W, 1), size(x, 2)), :matmul, lift(tr, x), lift(tr, W))
end
function NNlib.conv(x::AbstractArray{T,4}, w::DeviceLeaf{T,4}) where {T}
    tr = trace_of(w, x)
    tr === nothing && error("protograd_b200: a device model can only be called inside eval_with_pullback(f, m, :B200) or forward(m, x)")
    size(x, 3) == size(w, 3) || throw(DimensionMismatch("Conv expects $(size(w, 3)) channels, got $(size(x, 3))"))
    dims = (size(x, 1) - size(w, 1) + 1, size(x, 2) - size(w, 2) + 1, size(w, 4), size(x, 4))
    return Tracer{T,4}(tr, -1, dims, :conv, lift(tr, x), lift(tr, w))
end
# `.+ b` completes Linear / Conv
function broadcasted(::typeof(+), y::Tracer{T,N}, b::DeviceLeaf{T}) where {T,N}
    y.pending === :none && error("protograd_b200: only `W*x .+ b` / `conv(x, w) .+ b` add a parameter to a traced value")
    push_node!(y.trace, y.pending === :matmul ? OP_LINEAR : OP_CONV2D, (y.px, y.pw, lift(y.trace, b)), y.dims)
end
broadcasted(::typeof(NNlib.relu), x::Tracer) = push_node!(x.trace, OP_RELU, (x.id,), x.dims)
NNlib.softmax(x::Tracer; dims = 1) = (dims == 1 || error("softmax over the class axis only"); push_node!(x.trace, OP_SOFTMAX, (x.id,), x.dims))
function pool(kind, x::Tracer, k)
    k[1] == k[2] || error("square pooling windows only")
    dims = ((x.dims[1] - k[1]) ÷ k[1] + 1, (x.dims[2] - k[2]) ÷ k[2] + 1, x.dims[3], x.dims[4])
    push_node!(x.trace, kind, (x.id,), dims; k = k[1])
end
NNlib.maxpool(x::Tracer, k::NTuple{2,Int}) = pool(OP_MAXPOOL, x, k)
NNlib.meanpool(x::Tracer, k::NTuple{2,Int}) = pool(OP_MEANPOOL, x, k)
function ProtoGrad.dropout(x::Tracer, p; dims = :)
    ProtoGrad.is_training_mode() || return x
    dims === Colon() || error("Dropout with dims other than `:` is not on the B200 path")
    push_node!(x.trace, OP_DROPOUT, (x.id,), x.dims; p = Float64(p))
end
function ProtoGrad.cross_entropy(probs::Tracer{T}, label; dims = 1, agg = mean, global_batch = 0) where {T}
    (dims == 1 && agg === mean) || error("cross_entropy over the class axis with agg = mean only")
    push_node!(probs.trace, OP_LOSS_CE, (probs.id, lift(probs.trace, label)), (); denom = Float64(global_batch))
end
function ProtoGrad.mse(out::Tracer{T}, label; agg = mean, denom = 0) where {T}
    agg === mean || error("mse with agg = mean only")
    push_node!(out.trace, OP_LOSS_MSE, (out.id, lift(out.trace, label)), (); denom = Float64(denom))
end

# struct pg_op of the header (80 bytes)
struct PgOp
    kind::Int32
    dtype::Int32
    in0::Int32; in1::Int32; in2::Int32
    ndim::Int32
    s0::Int64; s1::Int64; s2::Int64; s3::Int64
    grad_leaf::Int32
    k::Int32
    p::Float64
    denom::Float64
end
function PgOp(n::Node)
    c = reverse(collect(Int64, n.dims))                       # Julia dims -> C order
    s = vcat(c, zeros(Int64, 4 - length(c)))
    i = vcat(Int32.(n.ins), fill(Int32(-1), 3 - length(n.ins)))
    PgOp(n.kind, n.dtype, i[1], i[2], i[3], length(c), s[1], s[2], s[3], s[4], n.grad_leaf, n.k, n.p, n.denom)
end

mutable struct Plan
    h::Ptr{Cvoid}
    inputs::Dict{Int,Ptr{Cvoid}}      # node id -> device staging buffer of a host input
end
const PLANS = Dict{Any,Plan}()
signature(tr::Trace, root::Int, math) = (math, root, [(n.kind, n.dtype, n.ins, n.dims, n.grad_leaf, n.k, n.p, n.denom) for n in tr.nodes])
function plan_for(tr::Trace, root::Int, math::Cint)
    get!(PLANS, signature(tr, root, math)) do
        ops = PgOp[PgOp(n) for n in tr.nodes]
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:pg_plan_create, LIB), Cint, (Ptr{Cvoid}, Ptr{PgOp}, Cint, Cint, Cint, Ref{Ptr{Cvoid}}), ctx().h, ops, length(ops), root, math, h))
        p = Plan(h[], Dict{Int,Ptr{Cvoid}}())
        finalizer(x -> ccall((:pg_plan_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), p)
        p
    end
end
"device pointers of every INPUT / PARAM node (host inputs are uploaded into per-plan staging buffers)"
function bind!(plan::Plan, tr::Trace)
    ptrs = fill(C_NULL, length(tr.nodes))
    for (i, n) in enumerate(tr.nodes)
        if n.kind == OP_PARAM
            ptrs[i] = ptr(n.data)
        elseif n.kind == OP_INPUT
            a = n.data
            buf = get!(plan.inputs, i) do
                p = Ref{Ptr{Cvoid}}(C_NULL)
                check(ccall((:pg_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx().h, sizeof(a), p))
                p[]
            end
            GC.@preserve a check(ccall((:pg_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, buf, pointer(a), sizeof(a)))
            ptrs[i] = buf
        end
    end
    sync()                                                    # pageable host sources may be temporaries
    return ptrs
end

const RNG = Ref((seed = UInt64(0x5EED5EED), step = UInt64(0)))
seed!(s::Integer) = (RNG[] = (seed = UInt64(s), step = UInt64(0)))

"""
    eval_with_pullback(f, dm::DeviceModel, :B200; math = :fp32)

The AD-backend seam of src/gradient.jl:1-2 / ext/ProtoGradZygoteExt.jl:6-13: the forward pass runs now and `out` is a
Number; `pb()` takes no arguments, runs the backward pass and returns a gradient with `typeof(grad) == typeof(dm)`
(test/test_models.jl:147-148) whose leaves are addressable (`Array(grad.W)`).
"""
function ProtoGrad.eval_with_pullback(f, dm::DeviceModel{M,T}, ::Val{:B200}; math::Symbol = :fp32, sample_offset::Integer = 0) where {M,T}
    tr = Trace{T}(Node[], IdDict{Any,Int}(), arena(dm).h)
    ACTIVE[] = tr
    out = try
        f(dm)
    finally
        ACTIVE[] = nothing
    end
    out isa Tracer{T,0} || error("the objective must return a scalar loss (cross_entropy / mse)")
    plan = plan_for(tr, out.id, math === :bf16 ? PG_MATH_BF16 : PG_MATH_FP32)
    ptrs = bind!(plan, tr)
    loss, gen = Ref{Ptr{Cvoid}}(C_NULL), Ref{Int64}(0)
    r = RNG[]
    RNG[] = (seed = r.seed, step = r.step + 1)
    check(ccall((:pg_plan_forward, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, UInt64, UInt64, Int64, Ref{Ptr{Cvoid}}, Ref{Int64}),
                plan.h, ptrs, C_NULL, r.seed, r.step, sample_offset, loss, gen))
    val = Ref{Cdouble}(0)
    check(ccall((:pg_loss_read, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ref{Cdouble}), ctx().h, code(T), loss[], val))
    function pullback(; on_ready = C_NULL, user = C_NULL)
        g = zero(dm)                                          # same type as the model (gradient.jl:4-13)
        gl = Ptr{Cvoid}[ptr(p) for p in allparams(g)]
        check(ccall((:pg_plan_backward, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), plan.h, gen[], gl, length(gl), on_ready, user))
        return g
    end
    return T(val[]), pullback
end

# the reference's Symbol entry point (src/gradient.jl:2) forwards no keyword arguments: this one does, for device models only
ProtoGrad.eval_with_pullback(f, dm::DeviceModel, backend::Symbol; kw...) = ProtoGrad.eval_with_pullback(f, dm, Val(backend); kw...)

"inference: `forward(dm, x)` == `m(x)` of the reference, host result"
function forward(dm::DeviceModel{M,T}, x; math::Symbol = :fp32) where {M,T}
    tr = Trace{T}(Node[], IdDict{Any,Int}(), C_NULL)
    ACTIVE[] = tr
    out = try
        dm(x)
    finally
        ACTIVE[] = nothing
    end
    plan = plan_for(tr, out.id, math === :bf16 ? PG_MATH_BF16 : PG_MATH_FP32)
    ptrs = bind!(plan, tr)
    check(ccall((:pg_plan_forward, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, UInt64, UInt64, Int64, Ptr{Ptr{Cvoid}}, Ptr{Int64}),
                plan.h, ptrs, C_NULL, 0, 0, 0, C_NULL, C_NULL))
    p, dt, pitch = Ref{Ptr{Cvoid}}(C_NULL), Ref{Cint}(0), Ref{Int64}(0)
    check(ccall((:pg_plan_value, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}, Ref{Cint}, Ref{Int64}), plan.h, out.id, p, dt, pitch))
    (dt[] == code(T) && pitch[] == out.dims[1]) || error("protograd_b200: forward() returns dense $(T) values only (the root of this trace is stored as dtype $(dt[]), pitch $(pitch[]))")
    res = Array{T}(undef, out.dims)
    GC.@preserve res check(ccall((:pg_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, pointer(res), p[], sizeof(res)))
    return res
end

# ---- the whole step of a small MLP in one cooperative launch (csrc/mlp_step.cu) -------------------------------------------
# struct pg_mlp_desc of the header (320 bytes); NTuples are inline storage, i.e. C arrays
struct PgMlpDesc
    n_layers::Int32
    opt::Int32
    batch::Int64
    dims::NTuple{5,Int64}
    x::Ptr{Cvoid}
    labels::Ptr{Cvoid}
    W::NTuple{4,Ptr{Cvoid}}
    b::NTuple{4,Ptr{Cvoid}}
    gW::NTuple{4,Ptr{Cvoid}}
    gb::NTuple{4,Ptr{Cvoid}}
    loss_out::Ptr{Cvoid}
    loss_scale::Float64
    arena_p::Ptr{Cvoid}
    arena_g::Ptr{Cvoid}
    arena_m::Ptr{Cvoid}
    arena_v::Ptr{Cvoid}
    arena_m_hat::Ptr{Cvoid}
    arena_v_hat::Ptr{Cvoid}
    arena_n::Int64
    stepsize::Float64
    beta1::Float64
    beta2::Float64
    eps::Float64
    it::Int64
    flags::Int32
end
pad4(v, fill) = ntuple(i -> i <= length(v) ? v[i] : fill, 4)

"""
    mega_step!(state, grad, x, label, loss) -> state

One iteration of the loop of examples/01_mnist_mlp.jl:44-55 — `eval_with_pullback(m -> cross_entropy(m(x), label), m)`,
`pb()`, `step!(state, grad)` — for a model that is Linear -> relu -> ... -> Linear -> softmax (2 to 4 Linear layers,
Float32, widths and batch <= 1536) as ONE launch.  `x` (in x batch), `label` (classes x batch) are device arrays of fixed
address, `grad` receives the gradient (same type as the model, test/test_models.jl:147-148), `loss` is a 1-element device
array; `state` is a GradientDescentState or an AdamState whose `variable` is the DeviceModel.
"""
function mega_step!(state, grad::DeviceModel{M,Float32}, x::Ptr{Cvoid}, label::Ptr{Cvoid}, loss::Ptr{Cvoid}, batch::Integer;
                    global_batch::Integer = batch) where {M}
    dm = state.variable
    ls, gs = allparams(inner(dm)), allparams(inner(grad))               # W1, b1, W2, b2, ... in vec(m) order
    L = length(ls) ÷ 2
    dims = Int64[size(ls[1], 2); [size(ls[2l - 1], 1) for l in 1:L]]    # Julia W[out, in]
    adam = state isa ProtoGrad.AdamState
    o = state.optimizer
    d = PgMlpDesc(L, adam ? 2 : 1, batch, ntuple(i -> i <= L + 1 ? dims[i] : 0, 5), x, label,
                  pad4([ptr(ls[2l - 1]) for l in 1:L], C_NULL), pad4([ptr(ls[2l]) for l in 1:L], C_NULL),
                  pad4([ptr(gs[2l - 1]) for l in 1:L], C_NULL), pad4([ptr(gs[2l]) for l in 1:L], C_NULL),
                  loss, 1.0 / global_batch, bp(dm), bp(grad), adam ? bp(state.m) : C_NULL, adam ? bp(state.v) : C_NULL,
                  adam ? bp(state.m_hat) : C_NULL, adam ? bp(state.v_hat) : C_NULL, np(grad), Float64(state.stepsize),
                  adam ? Float64(o.beta1) : 0.0, adam ? Float64(o.beta2) : 0.0, adam ? Float64(o.epsilon) : 0.0,
                  adam ? state.it : 1, isf64(state.stepsize))
    check(ccall((:pg_mlp_step, LIB), Cint, (Ptr{Cvoid}, Ref{PgMlpDesc}), ctx().h, Ref(d)))
    adam && (state.it += 1)                                             # adam.jl:40
    return state
end

# ---- data parallelism: one Julia process per GPU (MPI.jl or any launcher distributes the 128-byte id) ----------------
mutable struct DataParallel
    h::Ptr{Cvoid}
    rank::Int
    world::Int
end
function unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:pg_dp_unique_id, LIB), Cint, (Ptr{Cvoid},), id))
    return id
end
function DataParallel(rank::Integer, world::Integer, id::Vector{UInt8}; max_ctas::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pg_dp_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), ctx().h, rank, world, id, max_ctas, h))
    d = DataParallel(h[], rank, world)
    finalizer(x -> ccall((:pg_dp_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), d)
    return d
end
"in-place sum of the gradient arena over ranks; the compute stream is ordered behind it (asynchronously)"
function allreduce!(d::DataParallel, g::DeviceModel{M,T}) where {M,T}
    check(ccall((:pg_dp_allreduce, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint), d.h, bp(g), np(g), code(T)))
    check(ccall((:pg_dp_wait, LIB), Cint, (Ptr{Cvoid},), d.h))
    return g
end
function broadcast!(d::DataParallel, m::DeviceModel{M,T}; root::Integer = 0) where {M,T}
    check(ccall((:pg_dp_broadcast, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Cint), d.h, bp(m), np(m), code(T), root))
    check(ccall((:pg_dp_wait, LIB), Cint, (Ptr{Cvoid},), d.h))
    return m
end
"""
Overlapped exchange: buckets of finished gradient leaves are reduced while the remaining backward kernels run — the
library's own C callback is handed to the pullback, no Julia code runs in between:

    begin_buckets!(d, grad_arena_of_the_plan); g = pb(on_ready = bucket_callback(), user = d.h); finish_buckets!(d)
"""
bucket_callback() = cglobal((:pg_dp_bucket_ready, LIB))
begin_buckets!(d::DataParallel, g::DeviceModel; bucket_bytes::Integer = 24 << 20) =
    check(ccall((:pg_dp_bucket_begin, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), d.h, arena(g).h, bucket_bytes))
function finish_buckets!(d::DataParallel)
    n = Ref{Cint}(0)
    check(ccall((:pg_dp_bucket_finish, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}, Ptr{Int64}, Ptr{Int64}, Cint), d.h, n, C_NULL, C_NULL, 0))
    for k in 0:n[]-1
        check(ccall((:pg_dp_bucket_wait, LIB), Cint, (Ptr{Cvoid}, Cint), d.h, k))
    end
end

end # module

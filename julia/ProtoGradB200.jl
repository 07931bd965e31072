# ProtoGradB200.jl — Julia glue behind ProtoGrad's own extension seams (SOURCE ONLY: there is no
# Julia toolchain in the build image, so this file has never been executed; the Python host mirror
# in protograd.jl_b200/ drives exactly the same C ABI and is what the tests run).
#
# It would live next to ext/ProtoGradZygoteExt.jl as a second package extension and adds:
#   * DeviceLeaf{T,N}: an AbstractArray-like leaf type living at a fixed offset of one device arena
#     (src/layers.jl:15-18,38-41 keep their parametric struct definitions unchanged)
#   * eval_with_pullback(f, m, ::Val{:B200})           (src/gradient.jl:1-2 seam)
#   * Base.copyto!(dest::T, bc::Broadcasted{Style{T}})  for arena-backed models (src/model.jl:148)
#   * LinearAlgebra.dot / zero / copy / similar / vec   (src/model.jl:33,49-56,200-202)
#   * step!(state, grad) methods for the four optimizers (src/optimizers/*.jl)
module ProtoGradB200

using ProtoGrad
using ProtoGrad: Model, allparams
using LinearAlgebra

const LIB = joinpath(@__DIR__, "..", "protograd.jl_b200", "libprotograd_b200.so")

const PG_F32, PG_F64 = Cint(0), Cint(1)
code(::Type{Float32}) = PG_F32
code(::Type{Float64}) = PG_F64

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
ctx() = (isassigned(CTX) || (CTX[] = Context()); CTX[])

# ---- arena + leaves ---------------------------------------------------------------------------
mutable struct Arena{T}
    ptr::Ptr{Cvoid}
    n::Int                      # padded length (each leaf start padded to 32 elements)
end

function Arena{T}(n::Int) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pg_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx().h, n * sizeof(T), p))
    check(ccall((:pg_memset0, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, p[], n * sizeof(T)))
    a = Arena{T}(p[], n)
    finalizer(x -> ccall((:pg_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx().h, x.ptr), a)
    return a
end

struct DeviceLeaf{T,N} <: AbstractArray{T,N}
    arena::Arena{T}
    offset::Int                 # in elements
    dims::NTuple{N,Int}
end
Base.size(l::DeviceLeaf) = l.dims
ptr(l::DeviceLeaf{T}) where {T} = l.arena.ptr + l.offset * sizeof(T)
# scalar host access is a device round trip (correctness-only path, src/model.jl:79-99)
function Base.getindex(l::DeviceLeaf{T}, i::Int) where {T}
    out = Ref{T}()
    check(ccall((:pg_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, out, ptr(l) + (i - 1) * sizeof(T), sizeof(T)))
    return out[]
end
function Base.setindex!(l::DeviceLeaf{T}, v, i::Int) where {T}
    r = Ref{T}(convert(T, v))
    check(ccall((:pg_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, ptr(l) + (i - 1) * sizeof(T), r, sizeof(T)))
    check(ccall((:pg_sync, LIB), Cint, (Ptr{Cvoid},), ctx().h))
    return v
end

pad32(n) = cld(n, 32) * 32

"Move a CPU model onto the device: one arena in vec(m) order (src/model.jl:33), leaves become views."
function to_device(m::M) where {M<:Model}
    ps = allparams(m)
    T = eltype(first(ps))
    arena = Arena{T}(sum(pad32(length(p)) for p in ps))
    off = Ref(0)
    function move(a::AbstractArray)
        leaf = DeviceLeaf{T,ndims(a)}(arena, off[], size(a))
        GC.@preserve a check(ccall((:pg_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, ptr(leaf), pointer(a), sizeof(a)))
        off[] += pad32(length(a))
        return leaf
    end
    dm = ProtoGrad.map_recursively(move, m)       # src/model.jl:42-47
    check(ccall((:pg_sync, LIB), Cint, (Ptr{Cvoid},), ctx().h))
    return dm
end

arena_of(m::Model) = first(allparams(m)).arena
base_ptr(m::Model) = ptr(first(allparams(m)))
padded_length(m::Model) = sum(pad32(length(p)) for p in allparams(m))

# ---- model algebra: one kernel per broadcast statement (src/model.jl:148-152) ---------------------
isf64(s) = s isa Float64 ? Cint(1) : Cint(0)     # Julia promotion rule (SURVEY.md A.6)

"`dest .= a .- s .* g` and friends are pattern-matched on the flattened Broadcasted, as recursive_copyto! receives it."
function axpy!(dest::M, a::M, s::Number, g::M, sign::Integer) where {M<:Model}
    T = eltype(first(allparams(dest)))
    check(ccall((:pg_vec_axpy, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Cint, Int64),
                ctx().h, code(T), base_ptr(dest), base_ptr(a), Float64(s), base_ptr(g), sign, isf64(s), padded_length(dest)))
    return dest
end

function LinearAlgebra.dot(a::M, b::M) where {M<:Model}
    first(allparams(a)) isa DeviceLeaf || return invoke(dot, Tuple{Model,Model}, a, b)
    T = eltype(first(allparams(a)))
    r = Ref{Cdouble}(0)
    check(ccall((:pg_vec_dot, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Cdouble}),
                ctx().h, code(T), base_ptr(a), base_ptr(b), padded_length(a), r))
    return T(r[])
end

# ---- optimizers: step! as one fused kernel (src/optimizers/*.jl) ---------------------------------
function ProtoGrad.step!(state::ProtoGrad.GradientDescentState{S,V}, g::V) where {S,V<:Model}
    T = eltype(first(allparams(g)))
    check(ccall((:pg_opt_gd, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Cint, Ptr{Cvoid}),
                ctx().h, code(T), base_ptr(state.variable), base_ptr(g), padded_length(g), Float64(state.stepsize), isf64(state.stepsize), C_NULL))
    return state.variable
end

function ProtoGrad.step!(state::ProtoGrad.AdamState{R,V}, g::V) where {R,V<:Model}
    T = eltype(first(allparams(g)))
    o = state.optimizer
    check(ccall((:pg_opt_adam, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Cdouble, Cdouble, Cdouble, Int64, Cint, Ptr{Cvoid}),
                ctx().h, code(T), base_ptr(state.variable), base_ptr(state.m), base_ptr(state.v), base_ptr(state.m_hat), base_ptr(state.v_hat),
                base_ptr(g), padded_length(g), Float64(state.stepsize), Float64(o.beta1), Float64(o.beta2), Float64(o.epsilon), state.it, isf64(state.stepsize), C_NULL))
    return state.it += 1                              # adam.jl:40
end

# d = mu*d - s*g ; p += d                                  (src/optimizers/heavy_ball_method.jl:16-19)
function ProtoGrad.step!(state::ProtoGrad.HeavyBallMethodState{S,V}, g::V) where {S,V<:Model}
    T = eltype(first(allparams(g)))
    mu = state.optimizer.momentum
    flags = isf64(state.stepsize) | (mu isa Float64 ? Cint(2) : Cint(0))     # PG_SCALAR_F64 | PG_SCALAR2_F64
    check(ccall((:pg_opt_heavyball, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Cdouble, Cint, Ptr{Cvoid}),
                ctx().h, code(T), base_ptr(state.variable), base_ptr(state.step), base_ptr(g), padded_length(g),
                Float64(state.stepsize), Float64(mu), flags, C_NULL))
    return state.variable
end

# prev = p ; p -= s*g ; p += mu*(p - prev), mu popped from the stateful host-side iterator
# (src/optimizers/nesterov_method.jl:33-39; the sequence itself is :3-6)
function ProtoGrad.step!(state::ProtoGrad.NesterovState{S,V,M}, g::V) where {S,V<:Model,M}
    T = eltype(first(allparams(g)))
    mu = popfirst!(state.momentum_iterator)                                   # Float64
    check(ccall((:pg_opt_nesterov, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Cdouble, Cint, Ptr{Cvoid}),
                ctx().h, code(T), base_ptr(state.variable), base_ptr(state.variable_prev), base_ptr(g), padded_length(g),
                Float64(state.stepsize), Float64(mu), isf64(state.stepsize) | Cint(2), C_NULL))
    return state.variable
end

# ---- eval_with_pullback(f, m, ::Val{:B200}) ------------------------------------------------------
# f(m) is run once on Tracer values: `*`(::DeviceLeaf, ::Tracer), broadcast `.+`, relu., NNlib.conv,
# maxpool, reshape, softmax, cross_entropy and mse record ops into a plan (same scheme as
# protograd.jl_b200/trace.py); the plan fuses Linear+bias+ReLU, softmax+cross_entropy and ReLU' into
# the data-gradient epilogue, then calls pg_linear_fwd / pg_tc_linear_fwd / pg_conv2d_fwd /
# pg_tc_conv2d_fwd (when pg_tc_conv2d_supported) / pg_softmax_ce_fwd_bwd ... and, in pb(),
# pg_linear_bwd / pg_tc_linear_bwd / pg_conv2d_bwd / pg_tc_conv2d_bwd, writing
# dW/db straight into a gradient model built by `zero(m)` so that typeof(grad) == typeof(m)
# (test/test_models.jl:148).
function ProtoGrad.eval_with_pullback(f, m::Model, ::Val{:B200})
    plan, out = trace_and_compile(f, m)               # see protograd.jl_b200/trace.py for the algorithm
    run_forward!(plan)
    pb() = (g = gradient_container(m); run_backward!(plan, g); g)
    return out, pb
end

end # module

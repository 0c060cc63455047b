# ForwardDiffB200.jl -- the Julia side of the drop-in: a device array type whose methods are one
# `ccall` each into libfdb200.so (include/fdb200.h).
#
# NOT RUNNABLE IN THIS IMAGE (no `julia` binary): this file is the declarative statement of the
# bindings a ForwardDiff maintainer would add; the same C ABI is exercised by the Python/ctypes host
# in forwarddiff.jl_b200/fdb200/, which is what the tests and the benchmark run.  All logic that needs
# verification lives below the C ABI, and every C-ABI call SEQUENCE this file makes (argument order and
# types included) is executed by examples/julia_call_sequences.c on the GPU (tests/test_gpu_sharded.py),
# section by section under the same headings.
#
# Seam: ForwardDiff's config constructors call `similar(x, Dual{T,V,N})` (src/config.jl:122,159,
# 185-186,225-226).  For `x::B200Array` that returns a `B200DualArray{T,Float64,N}`; from then on
# Julia's dispatch routes seed!/extract_*/chunk_mode_* to the methods below.
module ForwardDiffB200

using ForwardDiff
using ForwardDiff: Dual, Partials, Chunk, GradientConfig, JacobianConfig, HessianConfig, DiffResults
import ForwardDiff: seed!, seed_zero_partials!, extract_gradient_chunk!, extract_jacobian_chunk!,
                    extract_jacobian!, extract_value!, chunk_mode_gradient!, chunk_mode_jacobian!, hessian!

const libfdb200 = joinpath(@__DIR__, "..", "lib", "libfdb200.so")

# ---- status -> Julia exceptions (fdb_status, include/fdb200.h) ---------------------------------
struct FdbError <: Exception; code::Cint; msg::String; end
function check(ctx::Ptr{Cvoid}, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:fdb_last_error, libfdb200), Cstring, (Ptr{Cvoid},), ctx))
    rc == 2 && throw(DimensionMismatch(msg))          # src/gradient.jl:46,88-91; src/jacobian.jl:89,164
    rc == 6 && throw(ArgumentError(msg))              # chunk size > structural_length, src/gradient.jl:116-118
    rc == 4 && throw(OutOfMemoryError())
    throw(FdbError(rc, msg))
end

# ---- context: one per process and GPU -------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:fdb_init, libfdb200), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r)
        rc == 0 || error(unsafe_string(ccall((:fdb_last_error, libfdb200), Cstring, (Ptr{Cvoid},), C_NULL)))
        c = new(r[]); finalizer(c -> ccall((:fdb_shutdown, libfdb200), Cint, (Ptr{Cvoid},), c.h), c); c
    end
end
const CTX = Ref{Context}()
context() = isassigned(CTX) ? CTX[] : (CTX[] = Context(parse(Int, get(ENV, "LOCAL_RANK", "0"))))

# ---- device arrays -----------------------------------------------------------------------------------
# level 0: Vector{Float64}; level 1: Vector{Dual{T,Float64,N}}; level 2: Vector{Dual{T,Dual{T,Float64,N},N}}
mutable struct B200Array{E} <: AbstractVector{E}
    h::Ptr{Cvoid}; n::Int; ctx::Context
end
dual_level(::Type{Float64}) = 0
dual_level(::Type{Dual{T,V,N}}) where {T,V,N} = 1 + dual_level(V)
chunk_of(::Type{Float64}) = 0
chunk_of(::Type{Dual{T,V,N}}) where {T,V,N} = N

function B200Array{E}(::UndefInitializer, n::Integer; ctx = context()) where {E}
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx.h, ccall((:fdb_array_alloc, libfdb200), Cint, (Ptr{Cvoid}, Int64, Cint, Cint, Ref{Ptr{Cvoid}}),
                       ctx.h, n, chunk_of(E), dual_level(E), r))
    a = B200Array{E}(r[], n, ctx)
    finalizer(a -> ccall((:fdb_array_free, libfdb200), Cint, (Ptr{Cvoid},), a.h), a)
end
function B200Array(x::Vector{Float64}; ctx = context())
    a = B200Array{Float64}(undef, length(x); ctx)
    GC.@preserve x check(ctx.h, ccall((:fdb_upload_values, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), a.h, x)); a
end
Base.size(a::B200Array) = (a.n,)
Base.Array(a::B200Array{Float64}) = (out = Vector{Float64}(undef, a.n);
    check(a.ctx.h, ccall((:fdb_download_values, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Float64}), a.h, out)); out)
Base.getindex(::B200Array, i...) = error("scalar indexing of a B200Array is not supported")
# THE seam: src/config.jl:122,159,185-186,225-226
Base.similar(a::B200Array, ::Type{D}) where {D<:Dual} = B200Array{D}(undef, a.n; ctx = a.ctx)
Base.similar(a::B200Array, ::Type{D}, dims::Dims) where {D} = B200Array{D}(undef, prod(dims); ctx = a.ctx)

# ---- sweep utilities (src/apiutils.jl:76-143; src/gradient.jl:74-81; src/jacobian.jl:95-118) ---------
function seed!(d::B200Array{Dual{T,V,N}}, x::B200Array{V}, index, seeds::NTuple{N,Partials{N,V}}, chunksize = N) where {T,V,N}
    check(d.ctx.h, ccall((:fdb_seed, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64), d.ctx.h, d.h, x.h, index, chunksize)); d
end
seed!(d::B200Array{Dual{T,V,N}}, x::B200Array{V}, seeds::NTuple{N,Partials{N,V}}) where {T,V,N} = seed!(d, x, 0, seeds, 0)
function seed_zero_partials!(d::B200Array{Dual{T,V,N}}, x::B200Array{V}, index = 0, count = (index == 0 ? 0 : N)) where {T,V,N}
    check(d.ctx.h, ccall((:fdb_seed_zero, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64), d.ctx.h, d.h, x.h, index, count)); d
end
function extract_gradient_chunk!(::Type{T}, result::B200Array, ydual::B200Array{<:Dual}, index, chunksize) where {T}
    check(result.ctx.h, ccall((:fdb_extract_gradient_chunk, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64),
                              result.ctx.h, result.h, ydual.h, index, chunksize)); result
end
function extract_jacobian_chunk!(::Type{T}, result::B200Array, ydual::B200Array{<:Dual}, index, chunksize) where {T}
    check(result.ctx.h, ccall((:fdb_extract_jacobian_chunk, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64),
                              result.ctx.h, result.h, length(ydual), ydual.h, index, chunksize)); result
end
function extract_jacobian!(::Type{T}, result::B200Array, ydual::B200Array{<:Dual}, n) where {T}
    check(result.ctx.h, ccall((:fdb_extract_jacobian, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64),
                              result.ctx.h, result.h, length(ydual), ydual.h, n)); result
end

# ---- array arithmetic: Base.broadcast / sum / prod on B200Array{<:Dual} ------------------------------
# op codes = fdb_op1 / fdb_op2 of include/fdb200.h (src/dual.jl:476-597,709-717)
const OP1 = Dict(:- => 1, :exp => 2, :log => 3, :sin => 4, :cos => 5, :tanh => 6, :sqrt => 7, :abs => 8, :abs2 => 9, :inv => 10,
                 :pow0 => 11, :pow1 => 12, :pow2 => 13, :pow3 => 14, :tan => 15, :exp2 => 16, :log2 => 17, :log10 => 18, :log1p => 19,
                 :expm1 => 20, :sinh => 21, :cosh => 22, :asin => 23, :acos => 24, :atan => 25, :cbrt => 26)
const OP2 = Dict(:+ => 1, :- => 2, :* => 3, :/ => 4, :^ => 5, :max => 6, :min => 7, :atan => 8, :hypot => 9)
function map1(op::Symbol, a::B200Array{D}) where {D<:Dual}
    out = similar(a, D)
    check(a.ctx.h, ccall((:fdb_map1, libfdb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), a.ctx.h, OP1[op], out.h, a.h)); out
end
function map2(op::Symbol, a::B200Array, b::B200Array)
    D = a isa B200Array{<:Dual} ? eltype(a) : eltype(b); out = similar(a, D)
    check(a.ctx.h, ccall((:fdb_map2, libfdb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), a.ctx.h, OP2[op], out.h, a.h, b.h)); out
end
function map2(op::Symbol, a::B200Array{D}, s::Real, scalar_first::Bool = false) where {D<:Dual}
    out = similar(a, D)
    check(a.ctx.h, ccall((:fdb_map2_scalar, libfdb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint),
                         a.ctx.h, OP2[op], out.h, a.h, Float64(s), scalar_first)); out
end
# A BroadcastStyle for B200Array would flatten a `Broadcasted` tree into these calls (or, next round,
# into one fused op-program launch); literal_pow maps `x .^ 2` to FDB_POW2 (src/dual.jl:587-597).
struct B200Style <: Base.Broadcast.AbstractArrayStyle{1} end
Base.BroadcastStyle(::Type{<:B200Array}) = B200Style()
function Base.sum(a::B200Array{D}) where {D<:Dual}
    out = B200Array{D}(undef, 1; ctx = a.ctx)
    check(a.ctx.h, ccall((:fdb_reduce, libfdb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), a.ctx.h, 1, out.h, a.h)); out
end
function Base.prod(a::B200Array{D}) where {D<:Dual}
    out = B200Array{D}(undef, 1; ctx = a.ctx)
    check(a.ctx.h, ccall((:fdb_reduce, libfdb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), a.ctx.h, 2, out.h, a.h)); out
end

# ---- whole-path drivers: catalogue targets run every chunk in one fused launch -------------------------
# (src/gradient.jl:114-167, src/jacobian.jl:169-244, src/hessian.jl:14-71)
catalogue_id(f) = nothing                 # users opt in:  ForwardDiffB200.catalogue_id(::typeof(rosenbrock)) = 1
function chunk_mode_gradient!(result::B200Array{Float64}, f::F, x::B200Array{Float64}, cfg::GradientConfig{T,V,N}) where {F,T,V,N}
    id = catalogue_id(f)
    id === nothing && return invoke(chunk_mode_gradient!, Tuple{Any,F,Any,GradientConfig{T,V,N}}, result, f, x, cfg)  # reference loop on device arrays
    rank, world = parse(Int, get(ENV, "RANK", "0")), parse(Int, get(ENV, "WORLD_SIZE", "1"))
    plan = Ref{NTuple{11,Int64}}()
    ccall((:fdb_plan_chunks, libfdb200), Cint, (Int64, Cint, Cint, Cint, Ptr{Cvoid}), length(x), N, rank, world, plan)
    value = B200Array{Float64}(undef, 1; ctx = x.ctx)
    check(x.ctx.h, ccall((:fdb_gradient_chunked, libfdb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                         x.ctx.h, id, x.h, N, plan[][8], plan[][9], result.h, value.h))
    return result                          # multi-rank: see sharded_gradient! below (peer-memory exchange) or all-gather the slices
end
# value.(ydual) -> y (src/apiutils.jl:9-12); the DiffResult form (apiutils.jl:5-7) calls this through DiffResults.value!
function extract_value!(::Type{T}, out::B200Array, ydual::B200Array{<:Dual}) where {T}
    check(out.ctx.h, ccall((:fdb_extract_value, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), out.ctx.h, out.h, ydual.h)); out
end

# array-valued catalogue targets carry their Float64 parameters: A (m x n column-major), b (m), alpha
struct JacobianTarget
    id::Int; m::Int
    A::Union{Nothing,B200Array{Float64}}; b::Union{Nothing,B200Array{Float64}}; alpha::Float64
end
jacobian_target(f) = nothing              # users opt in:  ForwardDiffB200.jacobian_target(f::MyLayer) = JacobianTarget(1, m, A, b, 0.0)
handle(a::Nothing) = C_NULL
handle(a::B200Array) = a.h
# chunk_mode_jacobian!(result, f, x, cfg) (src/jacobian.jl:218-223) and the f!(y, x) form (:225-244): `result` is the column-major
# m x n matrix stored flat; every chunk of this rank's range in one launch (DMMA contraction for tanh.(A*x .+ b))
function chunk_mode_jacobian!(result::B200Array{Float64}, f::F, x::B200Array{Float64}, cfg::JacobianConfig{T,V,N};
                              y::Union{Nothing,B200Array{Float64}} = nothing, lo = 0, hi = -1) where {F,T,V,N}
    t = jacobian_target(f)
    t === nothing && return invoke(chunk_mode_jacobian!, Tuple{Any,F,Any,JacobianConfig{T,V,N}}, result, f, x, cfg)   # reference loop on device arrays
    check(x.ctx.h, ccall((:fdb_jacobian_chunked, libfdb200), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Cdouble, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
        x.ctx.h, t.id, x.h, t.m, N, handle(t.A), t.m, handle(t.b), t.alpha, lo, hi, result.h, t.m, handle(y)))
    return result
end
chunk_mode_jacobian!(result::B200Array{Float64}, f!::F, y::B200Array{Float64}, x::B200Array{Float64}, cfg::JacobianConfig{T,V,N}) where {F,T,V,N} =
    chunk_mode_jacobian!(result, f!, x, cfg; y = y)
# jacobian!(J::Matrix, f, x::Vector, cfg) on host arrays -> fdb_jacobian_host (A, b: host arrays or nothing)
function jacobian_host!(J::Matrix{Float64}, y::Vector{Float64}, id::Integer, x::Vector{Float64}, ::Chunk{N}; A = nothing, b = nothing,
                        alpha = 0.0, reuse_params = false, ctx = context()) where {N}
    m = size(J, 1)
    GC.@preserve J y x A b check(ctx.h, ccall((:fdb_jacobian_host, libfdb200), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Int64, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Cdouble, Cint, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
        ctx.h, id, x, length(x), m, N, A === nothing ? C_NULL : pointer(A), m, b === nothing ? C_NULL : pointer(b), alpha, reuse_params, 0, -1, J, m, y))
    return J
end

# hessian!(result, f, x, cfg) (src/hessian.jl:31-37; the DiffResult form :66-71 reads value and gradient from the same sweep):
# nested Dual{T,Dual{T,V,N},N} sweeps of every (inner, outer) chunk pair in one launch
function hessian!(result::B200Array{Float64}, f::F, x::B200Array{Float64}, cfg::HessianConfig{T,V,N};
                  grad::Union{Nothing,B200Array{Float64}} = nothing, value::Union{Nothing,B200Array{Float64}} = nothing, lo = 0, hi = -1) where {F,T,V,N}
    id = catalogue_id(f)
    if id === nothing                                # a flattened f: fdb_program_hessian_chunked
        g = grad === nothing ? B200Array{Float64}(undef, length(x); ctx = x.ctx) : grad
        program_hessian!(result, g, trace_scalar(f, length(x)), x, Chunk{N}(); lo = lo, hi = hi)
        return result
    end
    check(x.ctx.h, ccall((:fdb_hessian_chunked, libfdb200), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
        x.ctx.h, id, x.h, N, lo, hi, result.h, length(x), handle(grad), handle(value)))
    return result
end
function hessian_host!(H::Matrix{Float64}, grad::Vector{Float64}, id::Integer, x::Vector{Float64}, ::Chunk{N}; ctx = context()) where {N}
    value = Ref{Float64}(0.0)
    GC.@preserve H grad x check(ctx.h, ccall((:fdb_hessian_host, libfdb200), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Cint, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ref{Float64}),
        ctx.h, id, x, length(x), N, 0, -1, H, size(H, 1), grad, value))
    return H, grad, value[]
end

# ---- ONE Julia session, all GPUs (fdb_multi_*): what gradient!/jacobian!/hessian! on host Vectors bind to when several devices
# are visible.  Device i of n sweeps the chunk range fdb_plan_chunks(xlen, N, i, n) and copies its slice into the host result. ----
mutable struct DeviceGroup
    h::Ptr{Cvoid}
    function DeviceGroup(devices::Union{Nothing,Vector{Cint}} = nothing)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        rc = devices === nothing ? ccall((:fdb_multi_init, libfdb200), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), 0, C_NULL, r) :
                                   ccall((:fdb_multi_init, libfdb200), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), length(devices), devices, r)
        rc == 0 || error(unsafe_string(ccall((:fdb_multi_last_error, libfdb200), Cstring, (Ptr{Cvoid},), C_NULL)))
        g = new(r[]); finalizer(g -> ccall((:fdb_multi_shutdown, libfdb200), Cint, (Ptr{Cvoid},), g.h), g); g
    end
end
function mcheck(g::DeviceGroup, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:fdb_multi_last_error, libfdb200), Cstring, (Ptr{Cvoid},), g.h))
    rc == 2 && throw(DimensionMismatch(msg)); rc == 6 && throw(ArgumentError(msg)); throw(FdbError(rc, msg))
end
function gradient_host!(g::DeviceGroup, out::Vector{Float64}, id::Integer, x::Vector{Float64}, ::Chunk{N}) where {N}
    value = Ref{Float64}(0.0)
    GC.@preserve out x mcheck(g, ccall((:fdb_multi_gradient_host, libfdb200), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Ref{Float64}), g.h, id, x, length(x), N, out, value))
    return out, value[]
end
function jacobian_host!(g::DeviceGroup, J::Matrix{Float64}, y::Vector{Float64}, id::Integer, x::Vector{Float64}, ::Chunk{N}; A = nothing,
                        b = nothing, alpha = 0.0, reuse_params = false) where {N}
    m = size(J, 1)
    GC.@preserve J y x A b mcheck(g, ccall((:fdb_multi_jacobian_host, libfdb200), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Int64, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Cdouble, Cint, Ptr{Float64}, Int64, Ptr{Float64}),
        g.h, id, x, length(x), m, N, A === nothing ? C_NULL : pointer(A), m, b === nothing ? C_NULL : pointer(b), alpha, reuse_params, J, m, y))
    return J
end
function hessian_host!(g::DeviceGroup, H::Matrix{Float64}, grad::Vector{Float64}, id::Integer, x::Vector{Float64}, ::Chunk{N}) where {N}
    value = Ref{Float64}(0.0)
    GC.@preserve H grad x mcheck(g, ccall((:fdb_multi_hessian_host, libfdb200), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Ref{Float64}), g.h, id, x, length(x), N, H, size(H, 1), grad, value))
    return H, grad, value[]
end

# Host Vectors: ForwardDiff.gradient!(out::Vector, f, x::Vector, cfg) for a catalogue f -> fdb_gradient_host
function gradient_host!(out::Vector{Float64}, f, x::Vector{Float64}, ::Chunk{N}; ctx = context()) where {N}
    value = Ref{Float64}(NaN)
    GC.@preserve out x check(ctx.h, ccall((:fdb_gradient_host, libfdb200), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Cint, Int64, Int64, Ptr{Float64}, Ref{Float64}),
        ctx.h, catalogue_id(f), x, length(x), N, 0, -1, out, value))
    return out, value[]
end

# ---- op-programs: a non-catalogue f is flattened from its Broadcasted tree into {op,dst,a,b} words and run as ONE
# batched launch (include/fdb200.h "op-programs"); Program mirrors fdb200/trace.py:Program --------------------
struct Program
    term::Matrix{Int32}; epi::Matrix{Int32}; consts::Vector{Float64}; n_acc::Int; halo::Int; result_reg::Int
end
# data arrays the flattened f closes over (program instruction kinds 7 / 8): bound for the following program_* calls
bind_params!(ctx::Context, arrays::Vector{B200Array{Float64}}) =
    check(ctx.h, ccall((:fdb_program_bind_params, libfdb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}), ctx.h, length(arrays), [a.h for a in arrays]))
function program_gradient!(result::B200Array{Float64}, p::Program, x::B200Array{Float64}, ::Chunk{N}; lo = 0, hi = -1) where {N}
    value = B200Array{Float64}(undef, 1; ctx = x.ctx)
    GC.@preserve p check(x.ctx.h, ccall((:fdb_program_gradient_chunked, libfdb200), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Int32}, Cint, Ptr{Float64}, Cint, Cint, Cint, Cint, Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
        x.ctx.h, p.term, size(p.term, 2), p.epi, size(p.epi, 2), p.consts, length(p.consts), p.n_acc, p.halo, p.result_reg,
        x.h, N, lo, hi, result.h, value.h))
    return result, value
end
function program_hessian!(H::B200Array{Float64}, grad::B200Array{Float64}, p::Program, x::B200Array{Float64}, ::Chunk{N}; lo = 0, hi = -1) where {N}
    value = B200Array{Float64}(undef, 1; ctx = x.ctx)
    GC.@preserve p check(x.ctx.h, ccall((:fdb_program_hessian_chunked, libfdb200), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Int32}, Cint, Ptr{Float64}, Cint, Cint, Cint, Cint, Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64,
         Ptr{Cvoid}, Ptr{Cvoid}),
        x.ctx.h, p.term, size(p.term, 2), p.epi, size(p.epi, 2), p.consts, length(p.consts), p.n_acc, p.halo, p.result_reg, x.h, N, lo, hi,
        H.h, length(x), grad.h, value.h))
    return H, grad, value
end

# ---- flattening a user f: the Julia twin of fdb200/trace.py (its basic form: one term part + scalar epilogue) ------------
# f is called once on a TraceArray; broadcasting over TraceArrays records nodes instead of computing (BroadcastStyle below),
# `sum` closes a reduction.  The recorded DAG is linearised into {op,dst,a,b} words exactly as trace.py:_emit does
# (post-order, registers released at last use).  Guards, data arrays and the two-pass form follow trace.py the same way.
const K_LOADX, K_UNARY, K_DD, K_DC, K_CD, K_ACC = 0, 1, 2, 3, 4, 5      # op codes: OP1 / OP2 above
mutable struct TNode; kind::Int; op::Int; args::Vector{TNode}; c::Float64; scalar::Bool; end
struct TraceArray <: AbstractVector{Float64}; node::TNode; n::Int; accs::Vector{TNode}; end      # a view x[a:b] or a broadcast result
struct TraceScalar; node::TNode; accs::Vector{TNode}; end                                       # sum(...) and scalar math on it
Base.size(t::TraceArray) = (t.n,)
Base.getindex(t::TraceArray, r::UnitRange) =                                                     # views directly on the input only
    TraceArray(TNode(K_LOADX, 0, TNode[], t.node.c + first(r) - 1, false), length(r), t.accs)
struct B200TraceStyle <: Broadcast.AbstractArrayStyle{1} end
Base.BroadcastStyle(::Type{TraceArray}) = B200TraceStyle()
B200TraceStyle(::Val{1}) = B200TraceStyle()
function Base.copy(bc::Broadcast.Broadcasted{B200TraceStyle})                                    # x .op y -> a node, nothing computed
    args = map(a -> a isa Broadcast.Broadcasted ? copy(a) : a, bc.args)
    arr = args[findfirst(a -> a isa TraceArray, args)]
    if bc.f === Base.literal_pow                    # x .^ 2 lowers to literal_pow.(^, x, Val(2)): FDB_POW0..3, src/dual.jl:587-597
        x, k = args[2], typeof(args[3]).parameters[1]
        (x isa TraceArray && k isa Integer && 0 <= k <= 3) || error("literal power outside 0..3 cannot be flattened")
        return TraceArray(TNode(K_UNARY, OP1[Symbol(:pow, k)], [x.node], 0.0, false), arr.n, arr.accs)
    end
    f = nameof(bc.f)
    if length(args) == 1
        return TraceArray(TNode(K_UNARY, OP1[f], [args[1].node], 0.0, false), arr.n, arr.accs)
    elseif args[1] isa TraceArray && args[2] isa TraceArray
        return TraceArray(TNode(K_DD, OP2[f], [args[1].node, args[2].node], 0.0, false), arr.n, arr.accs)
    elseif args[1] isa TraceArray                                                                # Dual (op) Real, src/dual.jl Real methods
        return TraceArray(TNode(K_DC, OP2[f], [args[1].node], Float64(args[2]), false), arr.n, arr.accs)
    else
        return TraceArray(TNode(K_CD, OP2[f], [args[2].node], Float64(args[1]), false), arr.n, arr.accs)
    end
end
function Base.sum(t::TraceArray)
    push!(t.accs, t.node)
    return TraceScalar(TNode(K_ACC, length(t.accs) - 1, TNode[], 0.0, true), t.accs)
end
for (f, code) in ((:+, 1), (:-, 2), (:*, 3), (:/, 4))
    @eval Base.$f(a::TraceScalar, b::TraceScalar) = TraceScalar(TNode(K_DD, $code, [a.node, b.node], 0.0, true), a.accs)
    @eval Base.$f(a::TraceScalar, b::Real) = TraceScalar(TNode(K_DC, $code, [a.node], Float64(b), true), a.accs)
    @eval Base.$f(a::Real, b::TraceScalar) = TraceScalar(TNode(K_CD, $code, [b.node], Float64(a), true), b.accs)
end
for (f, code) in ((:exp, 2), (:log, 3), (:sqrt, 7))
    @eval Base.$f(a::TraceScalar) = TraceScalar(TNode(K_UNARY, $code, [a.node], 0.0, true), a.accs)
end
# linearise the nodes reachable from `roots` with the given scalar-ness into instruction words (trace.py:_emit)
function emit(roots::Vector{TNode}, scalar::Bool, preloaded::Dict{TNode,Int}, consts::Vector{Float64})
    order = TNode[]; seen = Set{TNode}()
    visit(nd) = (nd in seen || nd.scalar != scalar) ? nothing : (push!(seen, nd); foreach(visit, nd.args); push!(order, nd))
    foreach(visit, roots)
    lastuse = Dict{TNode,Int}(); for (pos, nd) in enumerate(order), a in nd.args; lastuse[a] = pos; end
    reg = copy(preloaded); free = [r for r in 11:-1:0 if !(r in values(reg))]; words = Int32[]
    cidx(v) = (i = findfirst(==(v), consts); i === nothing ? (push!(consts, v); length(consts) - 1) : i - 1)
    for (pos, nd) in enumerate(order)
        haskey(reg, nd) && continue
        srcs = [reg[a] for a in nd.args]
        for a in nd.args; get(lastuse, a, 0) == pos && !haskey(preloaded, a) && push!(free, reg[a]); end
        dst = pop!(free); reg[nd] = dst
        w = nd.kind == K_LOADX ? (K_LOADX << 6, dst, 0, Int(nd.c)) :
            nd.kind == K_UNARY ? ((K_UNARY << 6) | nd.op, dst, srcs[1], 0) :
            nd.kind == K_DD ? ((K_DD << 6) | nd.op, dst, srcs[1], srcs[2]) :
            nd.kind == K_DC ? ((K_DC << 6) | nd.op, dst, srcs[1], cidx(nd.c)) : ((K_CD << 6) | nd.op, dst, cidx(nd.c), srcs[1])
        append!(words, Int32.(collect(w)))
    end
    return words, reg
end
function trace_scalar(f, n::Integer)
    accs = TNode[]
    y = f(TraceArray(TNode(K_LOADX, 0, TNode[], 0.0, false), n, accs))::TraceScalar
    consts = Float64[]
    term, reg = emit(accs, false, Dict{TNode,Int}(), consts)
    for (k, nd) in enumerate(accs); append!(term, Int32[K_ACC << 6, k - 1, reg[nd], 0]); end
    pre = Dict{TNode,Int}(); collectacc(nd) = (nd.kind == K_ACC && (pre[nd] = nd.op); foreach(collectacc, nd.args)); collectacc(y.node)
    epi, ereg = emit([y.node], true, pre, consts)
    halo = maximum(Int(term[4i]) for i in 1:length(term) ÷ 4 if term[4i - 3] >> 6 == K_LOADX; init = 0)
    return Program(reshape(term, 4, :), reshape(epi, 4, :), consts, length(accs), halo, ereg[y.node])
end
# ForwardDiff.gradient!(result, f, x::B200Array, cfg) for a non-catalogue f: flatten once, one batched launch
function chunk_mode_gradient_traced!(result::B200Array{Float64}, f::F, x::B200Array{Float64}, cfg::GradientConfig{T,V,N}) where {F,T,V,N}
    p = trace_scalar(f, length(x))
    return first(program_gradient!(result, p, x, Chunk{N}()))
end

# ---- the generic chunk loop as a replayed CUDA graph (fdb_graph.cu): chunk_mode_gradient! for a B200Array when f is an
# arbitrary Julia function of B200Array{<:Dual} (every array op inside f is a ccall recorded by stream capture) ----------
function chunk_mode_gradient_graph!(result::B200Array{Float64}, f::F, x::B200Array{Float64}, cfg::GradientConfig{T,V,N}) where {F,T,V,N}
    ctx, xdual, n = x.ctx, cfg.duals, length(x)
    seed!(xdual, x, 1, cfg.seeds); seed_zero_partials!(xdual, x, N + 1, n - N)        # first chunk eagerly, src/gradient.jl:133-138
    ydual = f(xdual); extract_gradient_chunk!(T, result, ydual, 1, N); seed_zero_partials!(xdual, x, 1)
    check(ctx.h, ccall((:fdb_cursor_set, libfdb200), Cint, (Ptr{Cvoid}, Int64, Int64), ctx.h, N + 1, min(N, n - N)))
    check(ctx.h, ccall((:fdb_graph_begin, libfdb200), Cint, (Ptr{Cvoid},), ctx.h))
    check(ctx.h, ccall((:fdb_seed_at_cursor, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ctx.h, xdual.h, x.h))
    ydual = f(xdual)                                                                   # temporaries must outlive the graph
    check(ctx.h, ccall((:fdb_extract_gradient_chunk_at_cursor, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ctx.h, result.h, ydual.h))
    check(ctx.h, ccall((:fdb_seed_zero_at_cursor, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ctx.h, xdual.h, x.h))
    check(ctx.h, ccall((:fdb_cursor_advance, libfdb200), Cint, (Ptr{Cvoid}, Int64, Int64), ctx.h, N, n))
    exec = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx.h, ccall((:fdb_graph_end, libfdb200), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), ctx.h, exec))
    nchunks = cld(n, N)
    check(ctx.h, ccall((:fdb_graph_launch, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), ctx.h, exec[], nchunks - 1))
    ccall((:fdb_graph_destroy, libfdb200), Cint, (Ptr{Cvoid},), exec[])
    return result
end

# ---- structured inputs (src/apiutils.jl:44-71): seed only structurally non-zero entries --------------------------------
const STRUCTURE = Dict(:dense => 0, :upper => 1, :lower => 2, :diagonal => 3)
seed_structured!(d::B200Array, x::B200Array, s::Symbol, rows, index, chunksize) =
    check(d.ctx.h, ccall((:fdb_seed_structured, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Int64, Int64),
                         d.ctx.h, d.h, x.h, STRUCTURE[s], rows, index, chunksize))
extract_gradient_chunk_structured!(res::B200Array, y::B200Array, s::Symbol, rows, index, chunksize) =
    check(res.ctx.h, ccall((:fdb_extract_gradient_chunk_structured, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Int64, Int64),
                           res.ctx.h, res.h, y.h, STRUCTURE[s], rows, index, chunksize))

# ---- multi-GPU (one Julia process per GPU, e.g. under MPI): results completed on every rank by the peer-memory exchange ----
# `allgather64(handle)::Vector{UInt8}` is the host's transport for the 64-byte IPC handles (e.g. MPI.Allgather).
mutable struct PeerExchange
    h::Ptr{Cvoid}
    ctx::Context
end
function PeerExchange(ctx::Context, ndoubles::Integer, rank::Integer, world::Integer, allgather64)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx.h, ccall((:fdb_exchange_create, libfdb200), Cint, (Ptr{Cvoid}, Int64, Cint, Cint, Ref{Ptr{Cvoid}}), ctx.h, ndoubles, rank, world, r))
    mine = Vector{UInt8}(undef, 64)
    check(ctx.h, ccall((:fdb_exchange_handle, libfdb200), Cint, (Ptr{Cvoid}, Ptr{UInt8}), r[], mine))
    all = allgather64(mine)                                                        # world * 64 bytes, rank order
    check(ctx.h, ccall((:fdb_exchange_connect, libfdb200), Cint, (Ptr{Cvoid}, Ptr{UInt8}), r[], all))
    return PeerExchange(r[], ctx)
end
# The same arena bound into an NVSwitch multicast object (csrc/fdb_multicast.cu): every result entry is stored once and the switch
# replicates it into all ranks' arenas.  `allgather16` / `bcast16` move the 16-byte {pid, fd} handles (MPI.Allgather / MPI.Bcast!),
# `hostbarrier()` is any barrier of the host transport (every device must be added before anyone binds, and every arena bound
# before anyone stores).  Throws (FDB_ERR_UNSUPPORTED) where the box cannot do it: fall back to PeerExchange.
function MulticastExchange(ctx::Context, ndoubles::Integer, rank::Integer, world::Integer, allgather16, bcast16, hostbarrier)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx.h, ccall((:fdb_exchange_create_multicast, libfdb200), Cint, (Ptr{Cvoid}, Int64, Cint, Cint, Ref{Ptr{Cvoid}}), ctx.h, ndoubles, rank, world, r))
    mine = Vector{UInt8}(undef, 16); mc = zeros(UInt8, 16)
    check(ctx.h, ccall((:fdb_exchange_share_handle, libfdb200), Cint, (Ptr{Cvoid}, Ptr{UInt8}), r[], mine))
    rank == 0 && check(ctx.h, ccall((:fdb_exchange_multicast_create, libfdb200), Cint, (Ptr{Cvoid}, Ptr{UInt8}), r[], mc))
    all = allgather16(mine); mc = bcast16(mc)
    check(ctx.h, ccall((:fdb_exchange_connect_multicast, libfdb200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt8}), r[], all, mc))
    hostbarrier()
    check(ctx.h, ccall((:fdb_exchange_multicast_bind, libfdb200), Cint, (Ptr{Cvoid},), r[]))
    hostbarrier()
    return PeerExchange(r[], ctx)
end
# opt-in: the catalogue gradient kernel sweeps `group` chunks per block and shares the chunk-invariant far terms (1 = independent sweeps)
set_chunk_grouping!(ctx::Context, group::Integer) = check(ctx.h, ccall((:fdb_set_chunk_grouping, libfdb200), Cint, (Ptr{Cvoid}, Cint), ctx.h, group))
function arena_view(x::PeerExchange, offset::Integer, n::Integer, ld::Integer = n)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(x.ctx.h, ccall((:fdb_exchange_view, libfdb200), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ref{Ptr{Cvoid}}), x.h, offset, n, ld, r))
    return B200Array{Float64}(r[], n, x.ctx)
end
barrier(x::PeerExchange) = check(x.ctx.h, ccall((:fdb_exchange_barrier, libfdb200), Cint, (Ptr{Cvoid},), x.h))
# chunk_mode_gradient! over this rank's chunk range; `result` and `value` are arena views, so on return (after the
# second barrier) every rank holds the whole gradient (src/gradient.jl:141-155 across ranks)
function sharded_gradient!(x::PeerExchange, result::B200Array{Float64}, value::B200Array{Float64}, fn_id, xs::B200Array{Float64}, ::Chunk{N},
                           rank, world) where {N}
    plan = Ref{NTuple{11,Int64}}()             # fdb_chunk_plan: fields 8, 9 = chunk_lo, chunk_hi
    ccall((:fdb_plan_chunks, libfdb200), Cint, (Int64, Cint, Cint, Cint, Ptr{Cvoid}), length(xs), N, rank, world, plan)
    barrier(x)                                                                     # peers have consumed the previous result
    check(x.ctx.h, ccall((:fdb_exchange_enable, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.h, x.h))
    try
        check(x.ctx.h, ccall((:fdb_gradient_chunked, libfdb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                             x.ctx.h, fn_id, xs.h, N, plan[][8], plan[][9], result.h, value.h))
    finally
        ccall((:fdb_exchange_enable, libfdb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.h, C_NULL)
    end
    barrier(x)                                                                     # every peer's stores have landed here
    return result
end

end # module

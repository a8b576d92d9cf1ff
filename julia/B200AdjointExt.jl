# B200AdjointExt.jl -- Julia-side binding of libb200adj.so (include/b200adj.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S ENVIRONMENT: Julia is not installed in the build image (SURVEY.md finding 6).
# The file is written against the same header the Python/ctypes host layer and the parity tests drive, and mirrors the
# extension precedent ext/SciMLSensitivityMooncakeExt.jl:123-240 of the reference: a new sensealg type plus a new
# method of SciMLBase._concrete_solve_adjoint returning (primal, pullback).  Everything the device path does not
# support (callbacks, mass matrices, structured parameters, unknown RHS families, off-grid save times) is delegated
# back to the wrapped reference sensealg.
module B200AdjointExt

using SciMLSensitivity, SciMLBase
using SciMLSensitivity: BacksolveAdjoint, InterpolatingAdjoint, QuadratureAdjoint, GaussAdjoint
import ChainRulesCore: NoTangent

const libb200adj = get(ENV, "B200ADJ_LIB", "libb200adj.so")

# ---- enums / cfg: field-for-field mirror of b200adj_cfg (168 bytes; checked against b200adj_sizeof_cfg) ----
@enum Family::Int32 FAM_LV = 0 FAM_LORENZ = 1 FAM_ROBERTSON = 2 FAM_SDE_LV = 3 FAM_MLP = 4 FAM_SDE_LINEAR = 5
const SA_CODE = Dict(InterpolatingAdjoint => Int32(0), GaussAdjoint => Int32(1), QuadratureAdjoint => Int32(2),
    BacksolveAdjoint => Int32(3), GaussKronrodAdjoint => Int32(4))
const ST_TSIT5_FIXED, ST_ROSENBROCK23, ST_EM, ST_EULER_HEUN, ST_TSIT5_ADAPTIVE = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const COST_EXPLICIT, COST_AFFINE = Int32(0), Int32(1)
const FLAG_NO_START, FLAG_NO_CHECKPOINTING, FLAG_CKPT_EVERY_STEP = UInt32(1), UInt32(2), UInt32(4)

"""
    B200PresetAffine(tstops, scale, shift; pscale = nothing, pshift = nothing)

The callback family the device path carries: at each `tstops[e]` the affect is `u .= scale[:, e] .* u .+ shift[:, e]` and,
optionally, `p .= pscale[:, e] .* p .+ pshift[:, e]` (`save_positions = (false, false)`).  It is a plain marker object:
passed as `callback =` to `solve(...; sensealg = B200Adjoint(...))` it is consumed by the method below; for the reference
path build the equivalent `PresetTimeCallback(tstops, affect!)`.
"""
struct B200PresetAffine
    tstops::Vector{Float64}; scale::Matrix{Float64}; shift::Matrix{Float64}
    pscale::Union{Nothing, Matrix{Float64}}; pshift::Union{Nothing, Matrix{Float64}}
end
B200PresetAffine(t, s, c; pscale = nothing, pshift = nothing) = B200PresetAffine(collect(Float64, t), s, c, pscale, pshift)

struct B200Cfg
    rhs_family::Int32; sensealg::Int32; stepper::Int32; dtype::Int32
    d::Int32; P::Int32; m::Int32; K::Int32
    N::Int64
    t0::Float64; t1::Float64; dt::Float64
    abstol::Float64; reltol::Float64
    quad_abstol::Float64; quad_reltol::Float64
    saveat::Ptr{Float64}
    shared_p::Int32; buffers_on_device::Int32; device::Int32; cost_kind::Int32
    cost_a::Float64; cost_b::Float64
    seed::UInt64; traj_offset::Int64
    checkpoint_every::Int32; flags::UInt32; mlp_hidden::Int32; block_threads::Int32
end

function __init__()
    sz = ccall((:b200adj_sizeof_cfg, libb200adj), UInt32, ())
    sz == sizeof(B200Cfg) || error("b200adj_cfg layout mismatch: C $sz vs Julia $(sizeof(B200Cfg))")
end

"""
    B200Adjoint(inner; family, block_threads = 0)

`sensealg = B200Adjoint(GaussAdjoint(); family = :lorenz)`: run the wrapped continuous adjoint on the B200 engine.
`family` names the hand-differentiated RHS family the problem's `f` belongs to (the role a user-supplied
`ODEFunction(f; vjp, vjp_p)` plays in src/derivative_wrappers.jl:284-359).
"""
struct B200Adjoint{Inner} <: SciMLSensitivity.AbstractAdjointSensitivityAlgorithm{0, true, Val{:central}}
    inner::Inner
    family::Family
    block_threads::Int32
end
B200Adjoint(inner; family::Symbol, block_threads = 0) =
    B200Adjoint(inner, getfield(@__MODULE__, Symbol("FAM_", uppercase(String(family)))), Int32(block_threads))

check(h, rc) = rc == 0 || error("b200adj error $rc: " *
    unsafe_string(ccall((:b200adj_last_error, libb200adj), Cstring, (Ptr{Cvoid},), h)))

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(cfg::B200Cfg)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:b200adj_create, libb200adj), Int32, (Ref{B200Cfg}, Ref{Ptr{Cvoid}}), cfg, ref)
        rc == 0 || error("b200adj_create failed ($rc): " *
            unsafe_string(ccall((:b200adj_last_error, libb200adj), Cstring, (Ptr{Cvoid},), C_NULL)))
        h = new(ref[])
        finalizer(x -> ccall((:b200adj_destroy, libb200adj), Int32, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

stepper_code(alg, adaptive = false) = nameof(typeof(alg)) === :Tsit5 ? (adaptive ? ST_TSIT5_ADAPTIVE : ST_TSIT5_FIXED) :
    nameof(typeof(alg)) === :EM ? ST_EM :
    nameof(typeof(alg)) === :EulerHeun ? ST_EULER_HEUN :
    nameof(typeof(alg)) === :Rosenbrock23 ? ST_ROSENBROCK23 : error("B200Adjoint: unsupported solver $(typeof(alg))")

# Ensemble entry: u0 is d x N (column i = member i, as materialised by prob_func on the host,
# test/Core4/ensembles.jl:22-24), p is a flat Vector (shared) or P x N matrix (per member).
function SciMLBase._concrete_solve_adjoint(
        prob::Union{SciMLBase.AbstractODEProblem, SciMLBase.AbstractSDEProblem},
        alg, sensealg::B200Adjoint, u0, p, originator::SciMLBase.ADOriginator, args...;
        save_start = true, save_end = true, saveat = eltype(prob.tspan)[], save_idxs = nothing,
        dt = nothing, kwargs...)
    # anything the device path does not cover goes to the reference implementation unchanged
    delegate() = SciMLBase._concrete_solve_adjoint(prob, alg, sensealg.inner, u0, p, originator, args...;
        save_start, save_end, saveat, save_idxs, kwargs...)
    prob.f.mass_matrix !== SciMLBase.I && return delegate()
    adaptive = get(kwargs, :adaptive, true) && !(alg isa Union{EM, EulerHeun})       # OrdinaryDiffEq default; fixed step needs dt
    # callbacks: the device carries preset-time affine affects (B200PresetAffine marks them, see INTEGRATION.md); anything
    # else -- continuous callbacks, state-dependent affects, extra saved points -- goes to the reference implementation
    events = nothing
    if haskey(kwargs, :callback)
        cb = kwargs[:callback]
        (cb isa B200PresetAffine && alg isa Tsit5 && adaptive) || return delegate()
        events = cb
    end
    (p isa AbstractVecOrMat{Float64} && u0 isa AbstractVecOrMat{Float64} && (adaptive || dt !== nothing)) || return delegate()

    t0, t1 = prob.tspan
    ts = saveat isa Number ? collect(t0:saveat:t1) : sort(collect(Float64, saveat))          # concrete_solve.jl:718-725,752-756
    save_start || (!isempty(ts) && ts[1] == t0 && popfirst!(ts))
    save_end || (!isempty(ts) && ts[end] == t1 && pop!(ts))
    U0 = Matrix{Float64}(permutedims(reshape(u0, size(u0, 1), :)))                           # -> [d][N] row-major = N x d column-major
    d, N = size(u0, 1), size(U0, 1)
    shared = p isa AbstractVector
    Pm = shared ? Vector{Float64}(p) : Matrix{Float64}(permutedims(p))
    P = shared ? length(p) : size(p, 1)
    flags = sensealg.inner isa BacksolveAdjoint && !sensealg.inner.checkpointing ? FLAG_NO_CHECKPOINTING : UInt32(0)
    (!save_start && !isempty(saveat) && t0 in saveat) && (flags |= FLAG_NO_START)               # concrete_solve.jl:962
    cfg = GC.@preserve ts B200Cfg(Int32(sensealg.family), SA_CODE[typeof(sensealg.inner).name.wrapper], stepper_code(alg, adaptive), 0,
        d, P, prob isa SciMLBase.AbstractSDEProblem ? d : 0, length(ts), N, t0, t1, something(dt, 0.0),
        get(kwargs, :abstol, 1e-6), get(kwargs, :reltol, 1e-3),
        sensealg.inner isa QuadratureAdjoint ? sensealg.inner.abstol : 1e-6,
        sensealg.inner isa QuadratureAdjoint ? sensealg.inner.reltol : 1e-3,
        pointer(ts), shared, 0, 0, COST_EXPLICIT, 0.0, 0.0, get(kwargs, :seed, UInt64(0)), 0, 1, flags, 0,
        sensealg.block_threads)
    h = try
        GC.@preserve ts Handle(cfg)
    catch
        return delegate()                                     # B200ADJ_ERR_UNSUPPORTED etc.
    end
    if events !== nothing        # [E], [E][d], [E][d], optional [E][P] x2 (row-major at the ABI = column-major transposed here)
        E = length(events.tstops)
        sc, sh = Matrix{Float64}(events.scale), Matrix{Float64}(events.shift)                 # d x E column-major = [E][d]
        ps = events.pscale === nothing ? C_NULL : pointer(Matrix{Float64}(events.pscale))
        pc = events.pshift === nothing ? C_NULL : pointer(Matrix{Float64}(events.pshift))
        GC.@preserve events sc sh check(h.ptr, ccall((:b200adj_set_events, libb200adj), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h.ptr, E, events.tstops, sc, sh, ps, pc))
    end
    K = length(ts)
    saved = Array{Float64}(undef, N, d, K)                    # [K][d][N] in C order
    status = Vector{Int32}(undef, N)
    GC.@preserve U0 Pm saved status check(h.ptr, ccall((:b200adj_forward, libb200adj), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}), h.ptr, U0, Pm, C_NULL, saved, status))
    u = [permutedims(saved[:, :, k]) for k in 1:K]            # d x N per save time
    out = SciMLBase.build_solution(prob, alg, ts, u; retcode = all(==(0), status) ? ReturnCode.Success : ReturnCode.Unstable)

    function b200_adjoint_backpass(Δ)
        Δa = Array{Float64}(undef, N, d, K)                   # accepts Matrix / VectorOfArray / Vector{Matrix} (concrete_solve.jl:777-868)
        for k in 1:K
            Δk = Δ isa AbstractArray{<:Number, 3} ? view(Δ, :, :, k) : Δ[k]
            Δa[:, :, k] .= permutedims(reshape(Δk, d, N))
        end
        du0 = Matrix{Float64}(undef, N, d)
        dp = shared ? Vector{Float64}(undef, P) : Matrix{Float64}(undef, N, P)
        GC.@preserve Δa du0 dp check(h.ptr, ccall((:b200adj_reverse, libb200adj), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, Δa, du0, dp))
        du0_out = reshape(permutedims(du0), size(u0))         # concrete_solve.jl:978
        dp_out = shared ? reshape(dp, size(p)) : permutedims(dp)   # :980-986
        if originator isa SciMLBase.TrackerOriginator || originator isa SciMLBase.ReverseDiffOriginator
            (NoTangent(), NoTangent(), du0_out, dp_out, NoTangent(), ntuple(_ -> NoTangent(), length(args))...)
        else
            (NoTangent(), NoTangent(), NoTangent(), du0_out, dp_out, NoTangent(), ntuple(_ -> NoTangent(), length(args))...)
        end
    end
    return out, b200_adjoint_backpass
end

end # module

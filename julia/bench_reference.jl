# bench_reference.jl -- the REAL reference arm of bench.py's headline workload (config C2), for a machine with Julia:
# SciMLSensitivity.jl's own GaussAdjoint / InterpolatingAdjoint over an EnsembleProblem on the host cores.
# NOT MEASURED IN THIS REPOSITORY'S ENVIRONMENT (Julia is not installed in the build image; bench.py --impl reference times
# the C restatement of the same algorithm, oracle/adjoint_oracle.c, and labels it kind = "port").
#
#   julia -t auto julia/bench_reference.jl [members = 8192] [steps = 5]
#
# Workload: Lorenz u0 = [1,0,0] + 0.1 z_i, p = [10, 28, 8/3] shared, T = 10, Tsit5 adaptive = false dt = 0.01, saveat = 0.1,
# loss = sum over save times of sum(abs2, u - 2)/2 (dgdu = u - 2, test/Core3/adjoint.jl:1169-1171); one "step" = the gradient
# of the summed ensemble loss wrt p and every u0_i.
using OrdinaryDiffEq, SciMLSensitivity, Zygote, Random, Statistics, Printf

function lorenz!(du, u, p, t)
    du[1] = p[1] * (u[2] - u[1]); du[2] = u[1] * (p[2] - u[3]) - u[2]; du[3] = u[1] * u[2] - p[3] * u[3]
    return nothing
end

function main(N = 8192, steps = 5; sensealg = GaussAdjoint(autojacvec = EnzymeVJP()))
    rng = Xoshiro(20260923)
    U0 = [1.0, 0.0, 0.0] .+ 0.1 .* randn(rng, 3, N)
    p = [10.0, 28.0, 8 / 3]
    ts = collect(0.0:0.1:10.0)
    dg(out, u, p, t, i) = (out .= u .- 2.0)
    function one_member(i)
        prob = ODEProblem(lorenz!, U0[:, i], (0.0, 10.0), p)
        sol = solve(prob, Tsit5(); adaptive = false, dt = 0.01, saveat = ts)
        du0, dp = adjoint_sensitivities(sol, Tsit5(); t = ts, dgdu_discrete = dg, sensealg, adaptive = false, dt = 0.01)
        return du0, vec(dp)
    end
    one_member(1)                                                   # compile
    times = Float64[]
    for _ in 1:steps
        t0 = time_ns()
        dps = Vector{Vector{Float64}}(undef, N)
        Threads.@threads for i in 1:N
            _, dps[i] = one_member(i)
        end
        dp = sum(dps)                                               # the reduction the outer AD performs (test/Core4/ensembles.jl:22-31)
        push!(times, (time_ns() - t0) / 1e9)
    end
    @printf("{\"impl\": \"reference (SciMLSensitivity.jl)\", \"metric\": \"ensemble adjoint trajectories/sec\", \"value\": %.1f, \"members\": %d, \"threads\": %d, \"s_per_step\": %.3f}\n",
        N / median(times), N, Threads.nthreads(), median(times))
end

main(length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 8192, length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 5)

# CPU figure of the REFERENCE itself for the bench.py workloads (BASELINE.md §4), for whoever has Julia and
# bat/EuclidianNormalizingFlows.jl installed.  The build image has neither, so bench.py times a C restatement of the
# reference's unfused algorithm instead (oracle/enf_ref_cpu.c, "kind": "port"); this script produces the number that
# restatement stands in for, on the same chain shapes, parameter distributions and synthetic data, and prints JSON
# lines in the shape of bench.py's cpu_baseline.
#
#   JULIA_NUM_THREADS=16 julia --project=/path/to/EuclidianNormalizingFlows.jl julia/bench_reference.jl [N]
#
# The reference is single-threaded apart from BLAS (src/householder_trafo.jl:4 is a broadcast + sum per reflection,
# no BLAS either), so the 1-thread figure is the faithful one; the threaded figure shards the columns over
# Threads.@threads the way bench.py's OpenMP port does.
using EuclidianNormalizingFlows, Random, Printf
using EuclidianNormalizingFlows: CenterStretch, CenterContract, JohnsonTrafo, ScaleShiftTrafo, HouseholderTrafo,
                                 with_logabsdet_jacobian, mvnormal_negll_trafograd

function c3_chain(rng, D = 16, K = 4, T = Float32)
    # tests/chains.py: spec ["hh4", "jo", "cs"] = CenterStretch ∘ JohnsonTrafo ∘ HouseholderTrafo
    hh = HouseholderTrafo(T.(randn(rng, D, K)))
    jo = JohnsonTrafo(T.(0.3 .* randn(rng, D)), T.(1 .+ 0.5 .* rand(rng, D)), T.(0.2 .* randn(rng, D)), T.(1 .+ 0.5 .* rand(rng, D)))
    cs = CenterStretch(T.(0.5 .+ rand(rng, D)), T.(0.5 .+ rand(rng, D)), T.(0.2 .* randn(rng, D)))
    cs ∘ jo ∘ hh
end

function c5_chain(rng, D = 32, K = 4, T = Float32)
    cc = CenterContract(T.(0.5 .+ rand(rng, D)), T.(0.5 .+ rand(rng, D)), T.(0.2 .* randn(rng, D)))
    jo = JohnsonTrafo(T.(0.3 .* randn(rng, D)), T.(1 .+ 0.5 .* rand(rng, D)), T.(0.2 .* randn(rng, D)), T.(1 .+ 0.5 .* rand(rng, D)))
    hh = HouseholderTrafo(T.(randn(rng, D, K)))
    ss = ScaleShiftTrafo(T.(0.5 .+ rand(rng, D)), T.(0.1 .* randn(rng, D)))
    ss ∘ hh ∘ jo ∘ cc
end

function timed(f, reps = 3)
    f()                                   # compile
    best = Inf
    for _ in 1:reps
        best = min(best, @elapsed f())
    end
    best
end

function threaded(fun, X, nt)
    N = size(X, 2)
    edges = round.(Int, range(0, N; length = nt + 1))
    Threads.@threads for t in 1:nt
        fun(view(X, :, edges[t]+1:edges[t+1]))
    end
end

function main()
    N = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 4_000_000
    rng = MersenneTwister(42)
    nt = Threads.nthreads()
    X3 = randn(rng, Float32, 16, N)
    f3 = c3_chain(rng)
    t1 = timed(() -> with_logabsdet_jacobian(f3, view(X3, :, 1:min(N, 1_000_000))))
    @printf("{\"leg\": \"C3 fwd+ladj\", \"value\": %.4g, \"unit\": \"samples/s\", \"cores\": 1, \"kind\": \"reference\"}\n", min(N, 1_000_000) / t1)
    tn = timed(() -> threaded(x -> with_logabsdet_jacobian(f3, x), X3, nt))
    @printf("{\"leg\": \"C3 fwd+ladj\", \"value\": %.4g, \"unit\": \"samples/s\", \"cores\": %d, \"kind\": \"reference\"}\n", N / tn, nt)
    n5 = min(N, 250_000)
    X5 = randn(rng, Float32, 32, n5)
    f5 = c5_chain(rng)
    tg = timed(() -> mvnormal_negll_trafograd(f5, X5))
    @printf("{\"leg\": \"C5 loss+gradient step (Zygote)\", \"value\": %.4g, \"unit\": \"samples/s\", \"cores\": 1, \"kind\": \"reference\"}\n", n5 / tg)
end

main()

# bench_argmax.jl -- time the REFERENCE's own argmax path (TwoSD.argmax_procedure, subprob.jl:141-169, and
# build_sasa_cut's accumulation loop, epigraph.jl:134-143) on the synthetic storm-shaped workload that
# `bench.py --instance synth128` uses, so the Julia number can sit next to the GPU number and to the C
# restatement that `bench.py` reports as `cpu_baseline` (SURVEY.md 8(d)).
#
# UNRUN: the build image has no Julia.  For anyone with Julia 1.9.3 and the reference checkout:
#
#     julia --project=/path/to/SQLP -t 1 julia/bench_argmax.jl /path/to/SQLP [K] [N] [s]
#
# The reference path is single-threaded (no Threads anywhere in src/), so `-t 1` is the honest setting;
# the script prints Threads.nthreads() with its result.  Inputs follow SURVEY.md 8(d) config C5:
# m2 = 512 with stochastic rows 0..s-1, n1 = 128, Tbar = one -1 per first-stage column on rows 128..255,
# rbar_j = 100 + 400 u(6, j), outcome tables rbar_j * {0.8, 0.9, 1.0, 1.1, 1.2}, pool pi_kj = 1000 (2 u(2, k m2 + j) - 1),
# x = 10 u(3, j), with u the splitmix64 counter generator shared by the oracle and the device.
using SparseArrays, LinearAlgebra, Printf

root = length(ARGS) >= 1 ? ARGS[1] : "."
include(joinpath(root, "src", "TwoSD.jl"))
using .TwoSD

K = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 1024
N = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 2000
s = length(ARGS) >= 4 ? parse(Int, ARGS[4]) : 128
const m2, n1 = 512, 128

function u01(seed::UInt64, idx::UInt64)::Float64
    z = seed ⊻ (idx * 0x9E3779B97F4A7C15)
    z += 0x9E3779B97F4A7C15
    z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9
    z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB
    z = z ⊻ (z >> 31)
    return Float64(z >> 11) / 9007199254740992.0
end
u(seed, idx) = u01(UInt64(seed), UInt64(idx))

rbar = spzeros(m2)
for j in 0:s-1
    rbar[j+1] = 100.0 + 400.0 * u(6, j)
end
Tbar = sparse(collect(129:256), collect(1:n1), fill(-1.0, n1), m2, n1)
row_lookup = Dict("R$(j)" => j + 1 for j in 0:m2-1)
col_lookup = Dict("X$(j)" => j + 1 for j in 0:n1-1)
coef = TwoSD.sdSubprobCoefficients(rbar, Tbar, spzeros(m2, 1), col_lookup, row_lookup)

dvs = TwoSD.sdDualVertexSet()
for k in 0:K-1
    push!(dvs, Float64[1000.0 * (2.0 * u(2, k * m2 + j) - 1.0) for j in 0:m2-1])
end

deltas = TwoSD.sdDeltaCoefficients[]
for i in 0:N-1
    scenario = TwoSD.spSmpsScenario([TwoSD.spSmpsPosition("RHS", "R$(j)") =>
                                     rbar[j+1] * (0.8 + 0.1 * floor(5.0 * u(1, i * s + j))) for j in 0:s-1])
    push!(deltas, TwoSD.delta_coefficients(coef, scenario))
end
x = Float64[10.0 * u(3, j) for j in 0:n1-1]

TwoSD.argmax_procedure(coef, deltas[1:min(N, 8)], x, dvs)                 # compile
t = @elapsed max_val, max_arg = TwoSD.argmax_procedure(coef, deltas, x, dvs)
evals = length(dvs) * N
@printf("{\"impl\": \"julia-reference\", \"metric\": \"scenario_x_vertex_argmax_evals_per_sec\", \"value\": %.6e, \"unit\": \"evals/s\", \"threads\": %d, \"K\": %d, \"N\": %d, \"s\": %d, \"m2\": %d, \"seconds\": %.4f, \"checksum\": %.17g}\n",
        evals / t, Threads.nthreads(), length(dvs), N, s, m2, t, sum(max_val))

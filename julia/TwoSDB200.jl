# TwoSDB200.jl -- reroutes TwoSD's argmax cut formation to libsqlp_b200.so (B200, sm_100a).
#
# UNRUN: Julia is not available in the build image.  This file is the reference-side binding a
# maintainer would add; `sqlp_b200/twosd.py` is its executable twin and is what the parity tests
# drive.  Usage, after `include("src/TwoSD.jl")`:
#
#     include("julia/TwoSDB200.jl"); using .TwoSDB200
#     TwoSDB200.enable!(TwoSD; lib = "/path/to/libsqlp_b200.so", device = 0)
#
# `enable!` overrides four methods of TwoSD, keeping their signatures and return types
# (SURVEY.md 8(b)); the fifth, the constructor, is left exactly as it is:
#     sdEpigraph(prob, w, lb)                  src/sd_algorithm/epigraph.jl:52-61   (unchanged: the device
#                                              epigraph is created LAZILY, at the first call below that
#                                              names the dual-vertex set -- the constructor does not)
#     add_scenario!(epi, scenario, weight)     src/sd_algorithm/epigraph.jl:81-96
#     Base.push!(dvs::sdDualVertexSet, v)      src/sd_algorithm/dual_set.jl:84-93
#     argmax_procedure(coef, deltas, x, dvs)   src/sd_algorithm/subprob.jl:141-169
#     build_sasa_cut(epi, x, dvs)::sdCut       src/sd_algorithm/epigraph.jl:125-146
# Host-side lists (scenario_list, scenario_weight, dvs.data) are still maintained so the rest of
# TwoSD (sd_iteration!, check_improvement, sync_cuts!) runs unchanged -- and so that a device
# epigraph bound late can replay the scenarios added before it existed.  No host code has to call
# anything new: `test/sd_test.jl` and `test/dual_set_test.jl` run as they are.
module TwoSDB200

using SparseArrays

const LIB = Ref{String}("libsqlp_b200.so")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

struct DeviceState
    pool::Ptr{Cvoid}
end
const POOLS = IdDict{Any,Ptr{Cvoid}}()        # sdDualVertexSet => sqlp_pool* (vertices of the set's main length)
const POOL_LEN = IdDict{Any,Int}()            # sdDualVertexSet => that length (fixed by the first vertex)
const SIDE = IdDict{Any,Dict{Int,Ptr{Cvoid}}}()   # sdDualVertexSet => pools of vectors of OTHER lengths
                                              # (dual_set.jl:26: never equal to the others, only counted)
const REPLAYED = IdDict{Any,Int}()            # sdEpigraph => scenarios of epi.scenario_list already on the device
const EPIS = IdDict{Any,Ptr{Cvoid}}()         # sdEpigraph      => sqlp_epi*
const DELTA_OWNER = IdDict{Any,Any}()         # epi.scenario_delta => epi
const TABLES = IdDict{Any,Vector{Any}}()      # sdSubprobCoefficients => position table

function check(status::Int32)
    status == 0 && return
    msg = unsafe_string(ccall((:sqlp_last_error, LIB[]), Cstring, ()))
    status == -4 && throw(UndefRefError())                      # what the reference throws
    error("libsqlp_b200 [$status]: $msg")
end

function context()
    if CTX[] == C_NULL
        out = Ref{Ptr{Cvoid}}(C_NULL)
        if N_GPUS[] > 1
            # this one Julia thread drives every GPU of the box (SURVEY.md 8(b)): pools replicated, epigraphs
            # sharded, one grouped all-gather per call -- nothing else in this file changes
            devs = Int32.(DEVICE[]:DEVICE[]+N_GPUS[]-1)
            check(ccall((:sqlp_ctx_create_multi, LIB[]), Int32, (Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}),
                        Int32(N_GPUS[]), devs, out))
        else
            check(ccall((:sqlp_ctx_create, LIB[]), Int32, (Int32, Ref{Ptr{Cvoid}}), Int32(DEVICE[]), out))
        end
        CTX[] = out[]
    end
    return CTX[]
end
const DEVICE = Ref{Int}(0)
const N_GPUS = Ref{Int}(1)

function new_pool(m2::Int)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sqlp_pool_create, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), context(), m2, out))
    return out[]
end

"The device pool holding the vertices of length `m2` of this set: the main pool (created on first use, with the
vertices pushed before it existed replayed in order) or, for any other length, a side pool of its own."
function pool_of(dvs, m2::Int)
    if !haskey(POOLS, dvs)
        POOL_LEN[dvs] = m2
        POOLS[dvs] = new_pool(m2)
        for dv in dvs.data
            length(dv.data) == m2 || continue
            ins = Ref{Int32}(0); idx = Ref{Int64}(0)
            check(ccall((:sqlp_pool_push, LIB[]), Int32,
                        (Ptr{Cvoid}, Ptr{Float64}, Ref{Int32}, Ref{Int64}), POOLS[dvs], dv.data, ins, idx))
        end
    end
    POOL_LEN[dvs] == m2 && return POOLS[dvs]
    return get!(() -> new_pool(m2), get!(() -> Dict{Int,Ptr{Cvoid}}(), SIDE, dvs), m2)
end

"Position table (row, col | -1) of the instance's random elements, fixed once from `sto`."
function position_table(T, coef, sto)
    rows = Int32[]; cols = Int32[]; order = Any[]
    for (pos, _) in sto.indep                       # Dict order is hash order: freeze it here
        push!(order, pos)
        push!(rows, Int32(coef.row_lookup[pos.row_name] - 1))
        push!(cols, (pos.col_name == "RHS" || pos.col_name == "rhs") ? Int32(-1) :
                    Int32(coef.col_lookup[pos.col_name] - 1))
    end
    return order, rows, cols
end

function enable!(T::Module; lib::String = "libsqlp_b200.so", device::Int = 0, n_gpus::Int = 1, sto = nothing)
    LIB[] = lib; DEVICE[] = device; N_GPUS[] = n_gpus
    sto === nothing && error("pass the spStoType so the position table can be resolved once")

    # --- the device half of an sdEpigraph, created at the first call that names the dual-vertex set ----
    # (sdEpigraph(prob, w, lb) itself -- epigraph.jl:52-61 -- is untouched: it does not know the set.)
    @eval T function device_epi(epi::sdEpigraph, dvs::sdDualVertexSet)
        if !haskey($EPIS, epi)
            coef = epi.subproblem_coef
            order, rows, cols = $position_table($T, coef, $sto)
            $TABLES[coef] = order
            r = coef.rhs; Tm = coef.transfer
            ridx = Int64.(r.nzind .- 1); rval = Float64.(r.nzval)
            colptr = Int64.(Tm.colptr .- 1); rowval = Int64.(Tm.rowval .- 1); nzval = Float64.(Tm.nzval)
            out = Ref{Ptr{Cvoid}}(C_NULL)
            $check(ccall((:sqlp_epi_create, $LIB[]), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Int64},
                 Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Int32}, Ref{Ptr{Cvoid}}),
                $context(), $pool_of(dvs, length(r)), length(r), size(Tm, 2), length(ridx), ridx, rval,
                colptr, rowval, nzval, length(rows), rows, cols, out))
            $EPIS[epi] = out[]
            $REPLAYED[epi] = 0
            $check(ccall((:sqlp_epi_set_weights, $LIB[]), Int32, (Ptr{Cvoid}, Float64, Float64), out[],
                         epi.objective_weight, epi.lower_bound))
            $DELTA_OWNER[epi.scenario_delta] = epi
            finalizer(e -> ccall((:sqlp_epi_destroy, $LIB[]), Int32, (Ptr{Cvoid},), $EPIS[e]), epi)
        end
        # scenarios added while no device epigraph existed (or through the host list alone): one batched call
        n0 = $REPLAYED[epi]; n1 = length(epi.scenario_list)
        if n1 > n0
            order = $TABLES[epi.subproblem_coef]
            vals = Matrix{Float64}(undef, length(order), n1 - n0)        # column-major = [n_new x s] row-major
            for (c, sc) in enumerate(epi.scenario_list[n0+1:n1])
                lookup = Dict(p => v for (p, v) in sc)
                vals[:, c] = Float64[lookup[p] for p in order]
            end
            $check(ccall((:sqlp_epi_add_scenarios, $LIB[]), Int32,
                         (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}), $EPIS[epi], n1 - n0, vals,
                         Float64.(epi.scenario_weight[n0+1:n1])))
            $REPLAYED[epi] = n1
        end
        return $EPIS[epi]
    end
    @eval T bind_device!(epi::sdEpigraph, dvs::sdDualVertexSet) = (device_epi(epi, dvs); epi)   # optional: bind early

    # --- add_scenario!(epi, scenario, weight) -- epigraph.jl:81-96 --------------------------------
    @eval T function add_scenario!(epi::sdEpigraph, scenario::spSmpsScenario, weight::Float64 = 1.0)
        push!(epi.scenario_list, scenario)
        push!(epi.scenario_weight, weight)
        epi.total_scenario_weight += weight
        # the deltas live on the device: epi.scenario_delta stays empty and only identifies the epigraph
        $DELTA_OWNER[epi.scenario_delta] = epi
        if haskey($EPIS, epi) && $REPLAYED[epi] == length(epi.scenario_list) - 1
            lookup = Dict(p => v for (p, v) in scenario)
            vals = Float64[lookup[p] for p in $TABLES[epi.subproblem_coef]]
            $check(ccall((:sqlp_epi_add_scenarios, $LIB[]), Int32,
                         (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}), $EPIS[epi], 1, vals, [weight]))
            $REPLAYED[epi] += 1
        end                                   # else: replayed when the device epigraph is bound (device_epi)
        return
    end

    # --- Base.push!(dvs, v) -- dual_set.jl:84-93 (returns the set) --------------------------------
    @eval T function Base.push!(dvs::sdDualVertexSet, new_vec::Vector{Float64})
        ins = Ref{Int32}(0); idx = Ref{Int64}(0)
        $check(ccall((:sqlp_pool_push, $LIB[]), Int32,
                     (Ptr{Cvoid}, Ptr{Float64}, Ref{Int32}, Ref{Int64}),
                     $pool_of(dvs, length(new_vec)), new_vec, ins, idx))
        # dvs.data keeps the reference's insertion order over ALL lengths (length, iterate: dual_set.jl:109-122);
        # the main pool's slots are the positions of the main-length vertices in it (what argmax_procedure maps back)
        ins[] == 1 && push!(dvs.data, sdDualVertex(new_vec))
        return dvs
    end

    # --- argmax_procedure(coef, delta_set, x, dvs; sense) -- subprob.jl:141-169 -------------------
    @eval T function argmax_procedure(coef::sdSubprobCoefficients, delta_set::Vector{sdDeltaCoefficients},
            x::Vector{Float64}, dual_vertices::sdDualVertexSet;
            sense::MOI.OptimizationSense = MIN_SENSE)
        epi = $DELTA_OWNER[delta_set]
        device_epi(epi, dual_vertices)
        n = length(epi.scenario_list)
        max_val = Vector{Float64}(undef, n); max_idx = Vector{Int64}(undef, n)
        $check(ccall((:sqlp_epi_argmax, $LIB[]), Int32,
                     (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Int64}),
                     $EPIS[epi], x, sense == MIN_SENSE ? Int32(0) : Int32(1), max_val, max_idx))
        any(<(0), max_idx) && throw(UndefRefError())
        m2 = $POOL_LEN[dual_vertices]
        main = [dv for dv in dual_vertices.data if length(dv.data) == m2]      # device slot k <-> k-th vertex of length m2
        max_arg = Ref{Vector{Float64}}[Ref(main[k + 1].data) for k in max_idx]
        return max_val, max_arg
    end

    # --- build_sasa_cut(epi, x, dvs)::sdCut -- epigraph.jl:125-146 --------------------------------
    @eval T function build_sasa_cut(epi::sdEpigraph, x::Vector{Float64}, dual_vertices::sdDualVertexSet)::sdCut
        alpha = Ref{Float64}(0); wm = Ref{Float64}(0); beta = zeros(length(x))
        $check(ccall((:sqlp_epi_build_cut, $LIB[]), Int32,
                     (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
                     device_epi(epi, dual_vertices), x, alpha, beta, wm, C_NULL))
        return sdCut(alpha[], beta, wm[])
    end

    # --- both cuts of one iteration (algorithm.jl:80,83) in one pass ------------------------------
    @eval T function build_two_cuts(epi::sdEpigraph, x_cand::Vector{Float64}, x_inc::Vector{Float64},
                                    dual_vertices::sdDualVertexSet)
        alpha = zeros(2); beta = zeros(length(x_cand), 2); wm = Ref{Float64}(0)
        $check(ccall((:sqlp_epi_build_cuts2, $LIB[]), Int32,
                     (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
                     device_epi(epi, dual_vertices), x_cand, x_inc, alpha, beta, wm, C_NULL))
        return sdCut(alpha[1], beta[:, 1], wm[]), sdCut(alpha[2], beta[:, 2], wm[])
    end
    # --- optional: the cut formation of one sd_iteration! in one call (algorithm.jl:45-55, 79-85) --
    # scenarios[i] is epigraph i's new scenario, duals the 2 * n_epi dual vertices found at the candidate
    # and the incumbent (in the order the reference pushes them).  Keeps the host-side lists in step.
    @eval T function sd_step!(cell::sdCell, scenarios::Vector{spSmpsScenario}, duals::Vector{Vector{Float64}})
        E = length(cell.epi)
        handles = Ptr{Cvoid}[device_epi(epi, cell.dual_vertices) for epi in cell.epi]
        vals = Float64[]
        for (epi, sc) in zip(cell.epi, scenarios)
            push!(epi.scenario_list, sc); push!(epi.scenario_weight, 1.0); epi.total_scenario_weight += 1.0
            $REPLAYED[epi] += 1
            lookup = Dict(p => v for (p, v) in sc)
            append!(vals, Float64[lookup[p] for p in $TABLES[epi.subproblem_coef]])
        end
        m2 = length(duals[1]); n1 = length(cell.x_candidate)
        V = reduce(vcat, duals)                                   # row-major [n_vertices x m2]
        ins = zeros(Int32, length(duals)); idx = zeros(Int64, length(duals))
        alpha = zeros(2, E); beta = zeros(n1, 2, E); wm = zeros(E)
        $check(ccall((:sqlp_cell_sd_step, $LIB[]), Int32,
                     (Int32, Ptr{Ptr{Cvoid}}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}, Ptr{Int64},
                      Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                     E, handles, vals, C_NULL, length(duals), V, ins, idx, cell.x_candidate, cell.x_incumbent,
                     alpha, beta, wm, C_NULL))
        for (v, d) in enumerate(duals)
            ins[v] == 1 && push!(cell.dual_vertices.data, sdDualVertex(d))
        end
        for (i, epi) in enumerate(cell.epi)
            push!(epi.cuts, sdCut(alpha[1, i], beta[:, 1, i], wm[i]))
            epi.incumbent_cut = sdCut(alpha[2, i], beta[:, 2, i], wm[i])
        end
        return
    end
    # --- optional: the cut list on the device (SURVEY.md 8(f) N1 / N3) ----------------------------
    # After build_two_cuts: take both cuts into the device list without a host round trip, ask the
    # device for the incumbent test, and fetch the master rows of sync_cuts! as one dense block.
    @eval T function commit_cuts!(epi::sdEpigraph; with_incumbent::Bool = true)
        $check(ccall((:sqlp_epi_cuts_commit, $LIB[]), Int32, (Ptr{Cvoid}, Int32), $EPIS[epi], Int32(with_incumbent)))
        return
    end
    @eval T function delete_cuts!(epi::sdEpigraph, delete_index::Vector{Int})      # deleteat!(epi.cuts, ...)
        idx = Int64.(delete_index .- 1)
        $check(ccall((:sqlp_epi_cuts_delete, $LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}), $EPIS[epi], length(idx), idx))
        return
    end
    @eval T function check_improvement_device(cell::sdCell, cost::Vector{Float64})
        handles = Ptr{Cvoid}[$EPIS[epi] for epi in cell.epi]
        out = zeros(4)
        $check(ccall((:sqlp_cell_check_improvement, $LIB[]), Int32,
                     (Int32, Ptr{Ptr{Cvoid}}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}),
                     length(handles), handles, cell.x_candidate, cell.x_incumbent, cost, INCUMBENT_SELECTION_Q, out))
        return sdImprovementInfo(out[1], out[2], out[3], out[4] != 0.0)
    end
    @eval T function master_rows(epi::sdEpigraph)      # rows of sync_cuts!: [alpha' beta'] per cut, incumbent last
        n = Ref{Int64}(0)
        $check(ccall((:sqlp_epi_master_rows, $LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}), $EPIS[epi], C_NULL, n))
        rows = zeros(length(epi.subproblem_coef.col_lookup) + 1, n[])      # column-major: one column per cut row
        n[] > 0 && $check(ccall((:sqlp_epi_master_rows, $LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}),
                                $EPIS[epi], rows, n))
        return rows
    end

    # --- rand(sto) on the device for a batch (smps_sto.jl:113-149; SURVEY.md 8(f) N2) -------------
    # Outcome tables / distribution parameters are uploaded once, in position-table order; the
    # scenarios never visit the host, so the host-side scenario_list is NOT extended here.
    @eval T function sample_scenarios!(epi::sdEpigraph, n_new::Int, seed::UInt64; weight_seed::UInt64 = UInt64(0))
        order = $TABLES[epi.subproblem_coef]
        s = length(order)
        dists = [$sto.indep[p] for p in order]
        kind = Int32[d isa spSmpsDiscreteDistribution ? 0 : d isa spSmpsNormalDistribution ? 1 : 2 for d in dists]
        a = Float64[k == 1 ? d.mean : k == 2 ? d.left : 0.0 for (k, d) in zip(kind, dists)]
        b = Float64[k == 1 ? d.variance : k == 2 ? d.right : 0.0 for (k, d) in zip(kind, dists)]
        mo = max(1, maximum(k == 0 ? length(d.value) : 1 for (k, d) in zip(kind, dists)))
        vals = zeros(mo, s); cdf = ones(mo, s); cnt = ones(Int32, s)      # column-major = [s][mo] row-major
        for (e, (k, d)) in enumerate(zip(kind, dists))
            k == 0 || continue
            cnt[e] = length(d.value); vals[1:cnt[e], e] = d.value; cdf[1:cnt[e], e] = cumsum(d.probability)
        end
        $check(ccall((:sqlp_epi_set_outcomes, $LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                     $EPIS[epi], mo, vals, cdf, cnt))
        $check(ccall((:sqlp_epi_set_distributions, $LIB[]), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
                     $EPIS[epi], kind, a, b))
        $check(ccall((:sqlp_epi_sample_scenarios, $LIB[]), Int32, (Ptr{Cvoid}, Int64, UInt64, UInt64),
                     $EPIS[epi], n_new, seed, weight_seed))
        tw = Ref{Float64}(0)
        $check(ccall((:sqlp_epi_counts, $LIB[]), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ref{Float64}),
                     $EPIS[epi], C_NULL, C_NULL, tw))
        epi.total_scenario_weight = tw[]
        return
    end
    # --- SMPS files read by the library (SURVEY.md 8(f) N4; smps_cor.jl / smps_tim.jl / smps_sto.jl,
    # smps_prob.jl:14-102, subprob.jl:15-69) -- the device epigraph is built straight from the three
    # files, so the O(m2 (n1 + n2)) normalized_coefficient sweep of extract_coefficients is not needed for
    # the device side.  The position table is the library's (order of first appearance in the .sto file).
    @eval T function bind_device_smps!(epi::sdEpigraph, dvs::sdDualVertexSet, cor_path::String, tim_path::String,
                                       sto_path::String)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        $check(ccall((:sqlp_smps_load, $LIB[]), Int32, (Cstring, Cstring, Cstring, Ref{Ptr{Cvoid}}),
                     cor_path, tim_path, sto_path, h))
        dims = zeros(Int64, 12)
        $check(ccall((:sqlp_smps_dims, $LIB[]), Int32, (Ptr{Cvoid}, Ptr{Int64}), h[], dims))
        m2 = dims[6]; s = dims[10]
        name(what, i) = begin
            buf = zeros(UInt8, 256)
            $check(ccall((:sqlp_smps_name, $LIB[]), Int32, (Ptr{Cvoid}, Int32, Int64, Ptr{UInt8}, Int64),
                         h[], Int32(what), i, buf, 256))
            unsafe_string(pointer(buf))
        end
        $TABLES[epi.subproblem_coef] = Any[spSmpsPosition(name(8, e), name(9, e)) for e in 0:s-1]
        out = Ref{Ptr{Cvoid}}(C_NULL)
        $check(ccall((:sqlp_epi_create_smps, $LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                     $context(), $pool_of(dvs, Int(m2)), h[], out))
        ccall((:sqlp_smps_destroy, $LIB[]), Int32, (Ptr{Cvoid},), h[])
        $EPIS[epi] = out[]
        $check(ccall((:sqlp_epi_set_weights, $LIB[]), Int32, (Ptr{Cvoid}, Float64, Float64), out[],
                     epi.objective_weight, epi.lower_bound))
        $DELTA_OWNER[epi.scenario_delta] = epi
        return epi
    end
    return nothing
end

end # module

/*
 * sqlp_b200.h -- C ABI of libsqlp_b200.so: the argmax cut-formation path of the
 * yhz0/SQLP `TwoSD` solver as hand-written sm_100a CUDA kernels.
 *
 * This is the drop-in boundary.  The reference has no FFI for this
 * path -- it is pure Julia -- so each entry point replaces a Julia METHOD, cited below
 * as file:line in the reference checkout; the `ccall` stubs a maintainer would add are
 * in INTEGRATION.md and julia/TwoSDB200.jl.
 *
 * Conventions
 *   - every function returns an int32 status: 0 = SQLP_OK, < 0 = error; the message of
 *     the calling thread's last error is sqlp_last_error().
 *   - plain pointers and sizes only; the caller owns every buffer it passes; the library
 *     copies inputs before returning and never retains a host pointer.
 *   - indices are 0-based (the Julia shim adds 1); matrices named [a x b] are row-major
 *     unless stated; Tbar is CSC as in SparseMatrixCSC but 0-based.
 *   - blocking calls return after the context's stream has drained.  `_dev` variants
 *     take DEVICE pointers, only enqueue work and do not synchronise.
 *   - a handle is driven by one host thread at a time.
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef SQLP_B200_H
#define SQLP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQLP_API __attribute__((visibility("default")))

typedef struct sqlp_ctx sqlp_ctx;   /* one GPU (+ its rank in a scenario-sharded job)   */
typedef struct sqlp_pool sqlp_pool; /* sdDualVertexSet, dual_set.jl:69-78               */
typedef struct sqlp_epi sqlp_epi;   /* the device half of sdEpigraph, epigraph.jl:17-45 */
typedef struct sqlp_smps sqlp_smps; /* spCorType + spTimType + spStoType + the stage-2 tables */

enum {
    SQLP_OK = 0,
    SQLP_E_INVALID = -1,     /* bad argument                                             */
    SQLP_E_CUDA = -2,        /* CUDA runtime error (no device, launch failure, ...)      */
    SQLP_E_UNSUPPORTED = -3, /* e.g. sense = MAX: the reference's MAX branch never
                                selects (subprob.jl:159-161)                             */
    SQLP_E_NO_ARGMAX = -4,   /* empty pool or all scores NaN/-Inf: the reference throws
                                UndefRefError at epigraph.jl:140                         */
    SQLP_E_NOMEM = -5,
    SQLP_E_NCCL = -6,
    SQLP_E_RANGE = -7,       /* scenario / vertex index out of range                     */
    SQLP_E_IO = -8           /* a file could not be opened                               */
};

enum { SQLP_MIN_SENSE = 0, SQLP_MAX_SENSE = 1 }; /* MOI.OptimizationSense */

SQLP_API const char *sqlp_version(void);
SQLP_API const char *sqlp_last_error(void);

/* Memory-safety check without a sanitizer (compute-sanitizer is closed on the GPU pool this was developed on): with
 * SQLP_GUARD=1 in the environment every device buffer of the library carries 512 bytes of a known pattern in front
 * of it and behind it; this call synchronises the device, reads all of them back and returns how many guarded
 * buffers are alive and how many guard zones were written to.  tests/test_gpu_guards.py. */
SQLP_API int32_t sqlp_guard_check(int64_t *buffers, int64_t *damaged);

/* ---------------------------------------------------------------- context ---------- */

/* One context per GPU.  `device` is the CUDA ordinal. */
SQLP_API int32_t sqlp_ctx_create(int32_t device, sqlp_ctx **out);

/* Scenario-sharded job, one process (or thread) per GPU.  `nccl_id` is the 128-byte
 * ncclUniqueId produced by sqlp_nccl_unique_id() on rank 0 and handed to every rank by
 * the host (MPI, torch.distributed, a file, ...).  Scenario g of an epigraph lives on
 * rank (g / 128) % world.  The pool is replicated. */
SQLP_API int32_t sqlp_nccl_unique_id(void *out128);
SQLP_API int32_t sqlp_ctx_create_dist(int32_t device, int32_t rank, int32_t world,
                                      const void *nccl_id, sqlp_ctx **out);
/* ONE host thread driving n GPUs -- what a single-threaded host such as the reference's sd_iteration!
 * (algorithm.jl:39-115) needs to reach every GPU of the box without becoming an SPMD job.  devices = NULL
 * means 0 .. n-1.  The returned context is used exactly like a single-GPU one: every handle created on it is
 * replicated (pools) or sharded (epigraphs: scenario g on GPU (g / 128) % n) behind the scenes, each call drives
 * all GPUs side by side on their own streams, the host hands pushed vertices to every GPU itself (no broadcast),
 * the per-epigraph partials of a call travel in one grouped ncclAllGather (ncclCommInitAll communicators) and
 * are summed in rank order on every GPU -- the same kernels and the same bits as one process per GPU.
 * sqlp_epi_argmax returns all scenarios in GLOBAL order.  Calls taking device pointers (`_dev`) are single-GPU
 * only and return SQLP_E_UNSUPPORTED here; the cut list (sqlp_epi_cuts_*), sqlp_epi_delta, sqlp_eval_dual and
 * the profiling calls act on the first GPU. */
SQLP_API int32_t sqlp_ctx_create_multi(int32_t n_gpus, const int32_t *devices, sqlp_ctx **out);
SQLP_API int32_t sqlp_ctx_destroy(sqlp_ctx *ctx);

/* Run on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the
 * context's own.  NULL restores the context's stream. */
SQLP_API int32_t sqlp_ctx_set_stream(sqlp_ctx *ctx, void *cuda_stream);
SQLP_API int32_t sqlp_ctx_synchronize(sqlp_ctx *ctx);
/* Number of kernels this context has launched so far. */
SQLP_API int32_t sqlp_ctx_launch_count(sqlp_ctx *ctx, int64_t *n);
/* CUDA-event stopwatch on the context's stream: start, stop (enqueue), elapsed (syncs). */
SQLP_API int32_t sqlp_ctx_timer_start(sqlp_ctx *ctx);
SQLP_API int32_t sqlp_ctx_timer_stop(sqlp_ctx *ctx);
SQLP_API int32_t sqlp_ctx_timer_elapsed_ms(sqlp_ctx *ctx, double *ms);
/* Accumulated device time (ms) and launch count of the contraction kernel since the last
 * call with reset != 0; measured with CUDA events around each launch when enabled. */
/* enable: 0 off, 1 every kernel class, otherwise a mask -- bit (1 + cls) turns on class cls of the list below
 * (2 = the contraction alone: two event records per launch instead of ten per SD iteration). */
SQLP_API int32_t sqlp_ctx_profile(sqlp_ctx *ctx, int32_t enable);
SQLP_API int32_t sqlp_ctx_profile_read(sqlp_ctx *ctx, int32_t reset, double *contract_ms,
                                       int64_t *contract_launches, double *contract_flops);
/* The same for every kernel class, arrays of SQLP_PROF_NCLASS = 8: 0 FP64 contraction (work = flops, counted with
 * the pool size each launch saw), 1 delta build, 2 cut reduction, 3 pool push, 4 bias vectors (work = algorithmic
 * bytes), 5 screening pass on tcgen05 (work = executed bf16 flops), 6 exact decision among its
 * candidates (work = scenario-points), 7 the FP64 sweep launched behind a screening pass (runs only on fall-back). */
#define SQLP_PROF_NCLASS 8
SQLP_API int32_t sqlp_ctx_profile_classes(sqlp_ctx *ctx, int32_t reset, double *ms /*[8]*/,
                                          int64_t *launches /*[8]*/, double *work /*[8]*/);
/* The argmax of subprob.jl:148-166 through a SCREENING pass: every score approximately on the 5th-generation
 * tensor cores (bf16 x 2 operands, fp32 accumulation, rigorous error bound), then the FP64 arithmetic of the
 * full sweep on the vertices that can still win -- same indices, same values, bit for bit.  mode 0: never
 * (every score in FP64), 1: automatic (default; large shapes, and not for a while after a pass fell back),
 * 2: whenever the shape allows it.  The environment variable SQLP_SCREEN sets the initial mode. */
SQLP_API int32_t sqlp_ctx_set_screen(sqlp_ctx *ctx, int32_t mode);

/* ---------------------------------------------------------------- dual-vertex pool -- */

/* sdDualVertexSet() -- dual_set.jl:76-78.  All vertices have length m2. */
SQLP_API int32_t sqlp_pool_create(sqlp_ctx *ctx, int64_t m2, sqlp_pool **out);
SQLP_API int32_t sqlp_pool_destroy(sqlp_pool *pool);

/* Base.push!(dvs, v) -- dual_set.jl:84-93, with the dedup rule of :24-53 reproduced bit
 * for bit (16 significant BINARY digits, sequential 1-norm hash as a gate, first match
 * wins, insertion order kept).  *index = slot of v or of the stored duplicate.
 * In a sharded job rank 0's vector is broadcast (NCCL) and every rank reaches the same
 * decision; `v` is ignored on the other ranks. */
SQLP_API int32_t sqlp_pool_push(sqlp_pool *pool, const double *v, int32_t *inserted,
                                int64_t *index);
/* n pushes in order (each sees the ones before it); one synchronisation at the end. */
SQLP_API int32_t sqlp_pool_push_batch(sqlp_pool *pool, int64_t n, const double *v /*[n x m2]*/,
                                      int32_t *inserted /*[n] or NULL*/,
                                      int64_t *index /*[n] or NULL*/);
/* Enqueue n pushes of device-resident vectors; no synchronisation, no readback. */
SQLP_API int32_t sqlp_pool_push_dev(sqlp_pool *pool, int64_t n, const double *d_v);
/* length(dvs) -- dual_set.jl:109-111 */
SQLP_API int32_t sqlp_pool_size(sqlp_pool *pool, int64_t *K);
/* dvs.data[index].data -- the stored vector (the reference returns Refs to these). */
SQLP_API int32_t sqlp_pool_get(sqlp_pool *pool, int64_t index, double *out /*[m2]*/);
/* hash_dual_vector(v) as computed on the device -- dual_set.jl:46-53 (debug/test). */
SQLP_API int32_t sqlp_pool_hash(sqlp_pool *pool, const double *v, uint64_t *hash);

/* ---------------------------------------------------------------- epigraph ---------- */

/* sdEpigraph(prob, w, lb) after extract_coefficients -- epigraph.jl:52-61,
 * subprob.jl:15-69.  rbar as (index, value) pairs, Tbar as 0-based CSC [m2 x n1], and the
 * stochastic-position table resolved once from the .sto file: element e perturbs
 * (pos_row[e], pos_col[e]); pos_col[e] = -1 means the RHS column (subprob.jl:113).
 * Positions must be distinct. */
SQLP_API int32_t sqlp_epi_create(sqlp_ctx *ctx, sqlp_pool *pool, int64_t m2, int64_t n1,
                                 int64_t r_nnz, const int64_t *r_idx, const double *r_val,
                                 const int64_t *T_colptr, const int64_t *T_rowval,
                                 const double *T_nzval, int64_t s, const int32_t *pos_row,
                                 const int32_t *pos_col, sqlp_epi **out);
SQLP_API int32_t sqlp_epi_destroy(sqlp_epi *epi);

/* add_scenario!(epi, scenario, weight) -- epigraph.jl:81-96 -- for n_new scenarios, with
 * delta_coefficients (subprob.jl:104-121) built on the device.  values[i][e] is the
 * realised value of table element e; weights NULL means 1.0 (algorithm.jl:46).
 * In a sharded job every rank passes the same arrays and keeps the scenarios it owns. */
SQLP_API int32_t sqlp_epi_add_scenarios(sqlp_epi *epi, int64_t n_new,
                                        const double *values /*[n_new x s]*/,
                                        const double *weights /*[n_new] or NULL*/);
SQLP_API int32_t sqlp_epi_add_scenarios_dev(sqlp_epi *epi, int64_t n_new, const double *d_values,
                                            const double *weights_host /*[n_new] or NULL*/);

/* Device-side sampling of INDEP DISCRETE elements (rand(sto), smps_sto.jl:113-149, with an
 * explicit counter generator instead of a global RNG): outcome tables are rectangular
 * [s x max_outcomes] with cnt[e] valid entries, cdf = running sum of probabilities.
 * Scenario g (0-based ordinal in this epigraph) takes, for element e,
 *   u = u01(seed, g*s + e),  value = vals[e][min(#{c : cdf[e][c] <= u}, cnt[e]-1)],
 * and weight 1.0 (weight_seed = 0) or 0.5 + u01(weight_seed, g).  u01 is the splitmix64
 * counter generator of SURVEY.md 8(d).  Each rank generates only the scenarios it owns. */
SQLP_API int32_t sqlp_epi_set_outcomes(sqlp_epi *epi, int64_t max_outcomes, const double *vals,
                                       const double *cdf, const int32_t *cnt);
/* Continuous INDEP elements (smps_sto.jl:118-127): kind[e] = 0 DISCRETE (outcome tables above),
 * 1 NORMAL(mean = a, variance = b), 2 UNIFORM(left = a, right = b).  A continuous element takes
 *   uo = (top 53 bits of the same counter stream + 1/2) / 2^53  in (0, 1),
 *   value = a + sqrt(b) * Phi^-1(uo)   |   a + (b - a) * uo      (one fma each).
 * Optional: without this call every element is DISCRETE. */
SQLP_API int32_t sqlp_epi_set_distributions(sqlp_epi *epi, const int32_t *kind /*[s]*/,
                                            const double *par_a /*[s]*/, const double *par_b /*[s]*/);
SQLP_API int32_t sqlp_epi_sample_scenarios(sqlp_epi *epi, int64_t n_new, uint64_t seed,
                                           uint64_t weight_seed);

/* Global scenario count, scenarios held by this rank, epi.total_scenario_weight. */
SQLP_API int32_t sqlp_epi_counts(sqlp_epi *epi, int64_t *n_global, int64_t *n_local,
                                 double *total_weight);

/* Score-equivalent vertices.  A vertex enters scores and cut coefficients only through the RELEVANT rows (a
 * random element lives there, or rbar != 0, or Tbar has an entry); elsewhere rbar - Tbar x is exactly zero for
 * every x.  Vertices equal on the relevant rows (a degenerate stage-2 LP returns many: storm's 16 384 harvested
 * duals are 3 599 classes) have bit-identical scores, and argmax_procedure keeps the FIRST maximum
 * (subprob.jl:156), so only the first vertex of a class can ever be selected: the sweep visits one column per
 * class and returns the same indices, values and cuts.  columns = classes of the current pool (= its size when
 * every row is relevant); SQLP_TWINS=0 in the environment turns the classification off. */
SQLP_API int32_t sqlp_epi_view_columns(sqlp_epi *epi, int64_t *columns, int64_t *relevant_rows);

/* delta_coefficients readback for LOCAL scenario i -- subprob.jl:104-121:
 * delta_rhs dense [m2], delta_T one value per table element (0 for RHS elements). */
SQLP_API int32_t sqlp_epi_delta(sqlp_epi *epi, int64_t local_scen, double *delta_rhs,
                                double *delta_T);

/* argmax_procedure(coef, delta_set, x, dual_vertices; sense) -- subprob.jl:141-169.
 * max_val[i], max_idx[i] (pool slot, -1 if nothing beat -Inf) for this rank's scenarios
 * in local order (n_local entries). */
SQLP_API int32_t sqlp_epi_argmax(sqlp_epi *epi, const double *x /*[n1]*/, int32_t sense,
                                 double *max_val, int64_t *max_idx);

/* build_sasa_cut(epi, x, dual_vertices)::sdCut -- epigraph.jl:125-146.
 * val (may be NULL) receives sum_i p_i max_val_i, the quantity of epigraph.jl:142. */
SQLP_API int32_t sqlp_epi_build_cut(sqlp_epi *epi, const double *x, double *alpha,
                                    double *beta /*[n1]*/, double *weight_mark, double *val);
/* The candidate cut and the regenerated incumbent cut of one iteration
 * (algorithm.jl:80 and :83) in one pass; when no random element touches Tbar the two
 * share a single contraction.  beta is two contiguous n1-vectors. */
SQLP_API int32_t sqlp_epi_build_cuts2(sqlp_epi *epi, const double *x_cand, const double *x_inc,
                                      double alpha[2], double *beta /*[2 x n1]*/,
                                      double *weight_mark, double *val /*[2] or NULL*/);
/* Same for every epigraph of a cell (the loop of algorithm.jl:79-85), one
 * synchronisation.  Outputs are [n_epi][2], [n_epi][2][n1], [n_epi]. */
SQLP_API int32_t sqlp_cell_build_cuts2(int32_t n_epi, sqlp_epi *const *epi, const double *x_cand,
                                       const double *x_inc, double *alpha, double *beta,
                                       double *weight_mark, double *val /*or NULL*/);
/* The cut formation of one SD iteration in ONE call and one synchronisation -- sd_iteration!,
 * algorithm.jl:45-55 and :79-85, minus the LP solves that produce its inputs: for every epigraph
 * add_scenario!(epi_e, scenario_e, weight_e); then push!(dual_vertices, v) for the n_vertices dual
 * vertices found at the candidate and the incumbent, in order; then both cuts of every epigraph.
 * values holds the epigraphs' value vectors back to back ([s_0 | s_1 | ...], table order), weights one
 * per epigraph (NULL = 1.0); inserted / index [n_vertices] as sqlp_pool_push (may be NULL); cut outputs
 * as sqlp_cell_build_cuts2.  The epigraphs share one context and one pool.  Same results, bit for bit,
 * as the separate calls; what it saves is their synchronisations and pageable copies. */
SQLP_API int32_t sqlp_cell_sd_step(int32_t n_epi, sqlp_epi *const *epi, const double *values,
                                   const double *weights, int64_t n_vertices, const double *vertices,
                                   int32_t *inserted, int64_t *index, const double *x_cand,
                                   const double *x_inc, double *alpha, double *beta,
                                   double *weight_mark, double *val /*or NULL*/);
/* Screening statistics of an epigraph (synchronises): out[0] passes seen by the host, [1] of which fell back to
 * the FP64 sweep, and for the last pass seen: [2] candidates emitted, [3] exact evaluations, [4] overflowed
 * candidate lists, [5] bit 0: non-finite / out-of-range operands, >> 1: passes that succeeded but needed more than
 * eight exact evaluations per scenario-point (the automatic mode then leaves the pass out for a while: the FP64
 * sweep is quicker on such pools), [6..7] vertices that could win at each point. */
SQLP_API int32_t sqlp_epi_screen_stats(sqlp_epi *epi, int64_t *out /*[8]*/);
/* Enqueue only: d_x2 = [x_cand | x_inc] on the device, d_out = [2][n1 + 2] on the device
 * holding (alpha, beta[n1], val) per x.  Errors such as a missing argmax surface at the
 * next blocking call. */
SQLP_API int32_t sqlp_epi_build_cuts2_dev(sqlp_epi *epi, const double *d_x2, double *d_out);
/* The same for every epigraph of a cell in one call: d_out = [n_epi][2][n1 + 2].  Epigraphs with the same
 * template share their bias vectors; in a sharded job the whole cell costs ONE all-gather. */
SQLP_API int32_t sqlp_cell_build_cuts2_dev(int32_t n_epi, sqlp_epi *const *epi, const double *d_x2,
                                           double *d_out);

/* ---------------------------------------------------------------- cut list (next rows N1, N3) --- */

/* The epigraph's cuts kept on the device, so that the cuts the reduction just produced reach the
 * incumbent test and the master without a round trip each.  A cut is (alpha, beta[n1], weight_mark),
 * never scaled (sdCut, epigraph.jl:5-12).  In a sharded job every rank holds the same list. */

/* epi.objective_weight, epi.lower_bound -- epigraph.jl:27-31 (defaults 1.0, 0.0). */
SQLP_API int32_t sqlp_epi_set_weights(sqlp_epi *epi, double objective_weight, double lower_bound);
/* push!(epi.cuts, sdCut(alpha, beta, weight_mark)) with host values. */
SQLP_API int32_t sqlp_epi_cuts_push(sqlp_epi *epi, double alpha, const double *beta, double weight_mark);
/* epi.incumbent_cut = sdCut(...); beta = NULL means `nothing`. */
SQLP_API int32_t sqlp_epi_cuts_set_incumbent(sqlp_epi *epi, double alpha, const double *beta,
                                             double weight_mark);
/* algorithm.jl:76-84 after a cut formation: snapshot the list (sdEpigraphInfo, f_{k-1}), then
 * push!(epi.cuts, candidate cut) and, if with_incumbent, epi.incumbent_cut = incumbent cut -- both
 * taken on the device from the last sqlp_epi_build_cuts2 / sqlp_cell_build_cuts2 /
 * sqlp_epi_build_cut result, with weight_mark = total_scenario_weight. */
SQLP_API int32_t sqlp_epi_cuts_commit(sqlp_epi *epi, int32_t with_incumbent);
/* deleteat!(epi.cuts, idx) -- algorithm.jl:69; idx ascending, 0-based. */
SQLP_API int32_t sqlp_epi_cuts_delete(sqlp_epi *epi, int64_t n, const int64_t *idx);
SQLP_API int32_t sqlp_epi_cuts_count(sqlp_epi *epi, int64_t *n_cuts, int32_t *has_incumbent);
/* Read one cut back; index -1 is the incumbent cut. */
SQLP_API int32_t sqlp_epi_cuts_get(sqlp_epi *epi, int64_t index, double *alpha, double *beta /*[n1]*/,
                                   double *weight_mark);
/* evaluate_epigraph(epi | info, x) including the epigraph's weight -- epigraph.jl:177-220.
 * which = 0: the current list; 1: the snapshot taken by the last commit. */
SQLP_API int32_t sqlp_epi_evaluate(sqlp_epi *epi, const double *x /*[n1]*/, int32_t which, double *out);
/* The rows sync_cuts! adds to the master (cell.jl:163-202, add_cut_to_master! epigraph.jl:101-117):
 * rows[j] = (discount alpha + (1 - discount) lb, discount beta[n1]), discount = weight_mark /
 * total_scenario_weight, the incumbent cut last and undiscounted; one dense [n_rows x (1 + n1)]
 * block in one copy.  rows = NULL only returns the count. */
SQLP_API int32_t sqlp_epi_master_rows(sqlp_epi *epi, double *rows, int64_t *n_rows);
/* check_improvement(f_last, f_current, x_cand, x_inc, ...) -- improvement.jl:19-49 -- over the
 * epigraphs of a cell; cost[n1] is the linear first-stage objective (cell.objf_original),
 * q_factor the reference's INCUMBENT_SELECTION_Q = 0.2.  out4 = {candidate_estimation,
 * incumbent_estimation, required_improvement, is_improved (0 | 1)}. */
SQLP_API int32_t sqlp_cell_check_improvement(int32_t n_epi, sqlp_epi *const *epi, const double *x_cand,
                                             const double *x_inc, const double *cost, double q_factor,
                                             double *out4);

/* ---------------------------------------------------------------- SMPS reader (next row N4) ------ */

/* read_cor (smps_cor.jl:26-194), read_tim (smps_tim.jl:30-64), read_sto (smps_sto.jl:41-111;
 * sto_path NULL or "" = no .sto file), the stage-2 split of get_smps_stage_template
 * (smps_prob.jl:14-102) and extract_coefficients (subprob.jl:15-69) in native host code: the flat
 * tables the device path consumes, without JuMP.  Host-only, needs no GPU.  Format rules kept from
 * the reference: '*' comment lines, a line starting in column 1 is a section header, unsupported
 * sections / bound types / INDEP keywords are errors, the first row must be the 'N' row, later
 * COLUMNS entries overwrite earlier ones, defaults 0 <= x < +Inf, exact zeros are not stored in
 * Tbar / W / rbar.  The reference holds the random elements in a Dict (hash order); here element
 * e is the e-th DISTINCT (column, row) position in order of first appearance in the .sto file. */
SQLP_API int32_t sqlp_smps_load(const char *cor_path, const char *tim_path, const char *sto_path,
                                sqlp_smps **out);
SQLP_API int32_t sqlp_smps_destroy(sqlp_smps *smps);

enum { /* indices into dims[] of sqlp_smps_dims */
    SQLP_SMPS_ROWS = 0,      /* rows of the cor file, objective row included               */
    SQLP_SMPS_COLS = 1,
    SQLP_SMPS_COR_NNZ = 2,   /* stored (row, column) entries as written, zeros included    */
    SQLP_SMPS_N1 = 3,        /* first-stage columns = columns of Tbar                      */
    SQLP_SMPS_N2 = 4,
    SQLP_SMPS_M2 = 5,        /* stage-2 rows = length of a dual vertex                     */
    SQLP_SMPS_T_NNZ = 6,
    SQLP_SMPS_W_NNZ = 7,
    SQLP_SMPS_R_NNZ = 8,
    SQLP_SMPS_S = 9,         /* random elements                                            */
    SQLP_SMPS_MAX_OUTCOMES = 10,
    SQLP_SMPS_PERIODS = 11,
    SQLP_SMPS_NDIMS = 12
};
SQLP_API int32_t sqlp_smps_dims(sqlp_smps *smps, int64_t *dims /*[SQLP_SMPS_NDIMS]*/);

enum { /* `what` of sqlp_smps_name */
    SQLP_SMPS_COR_NAME = 0, SQLP_SMPS_TIM_NAME = 1, SQLP_SMPS_STO_NAME = 2,
    SQLP_SMPS_ROW_NAME = 3, SQLP_SMPS_COL_NAME = 4,                /* cor.row_names / col_names */
    SQLP_SMPS_PERIOD_NAME = 5, SQLP_SMPS_PERIOD_COL = 6, SQLP_SMPS_PERIOD_ROW = 7,
    SQLP_SMPS_ELEM_COL = 8, SQLP_SMPS_ELEM_ROW = 9                 /* spSmpsPosition of element e */
};
/* NUL-terminated copy of a name; SQLP_E_RANGE if index is out of range or buf too small. */
SQLP_API int32_t sqlp_smps_name(sqlp_smps *smps, int32_t what, int64_t index, char *buf,
                                int64_t buflen);
/* spCorType: directions[rows] ('N','G','L','E'), rhs[rows], bounds[cols] and template_matrix as
 * 0-based CSC with COR_NNZ entries (rows ascending per column).  Any pointer may be NULL. */
SQLP_API int32_t sqlp_smps_cor(sqlp_smps *smps, char *directions, double *rhs, double *lower,
                               double *upper, int64_t *colptr /*[cols+1]*/, int64_t *rowval,
                               double *nzval);
/* sdSubprobCoefficients (subprob.jl:4-13): rbar dense [m2], Tbar CSC [m2 x n1], W CSC [m2 x n2],
 * plus the stage-2 cost row [n2] and the first-stage cost row [n1].  Any pointer may be NULL. */
SQLP_API int32_t sqlp_smps_stage2(sqlp_smps *smps, double *rbar, int64_t *T_colptr,
                                  int64_t *T_rowval, double *T_nzval, int64_t *W_colptr,
                                  int64_t *W_rowval, double *W_nzval, double *cost, double *x_cost);
/* The stochastic-position table and the distributions, arrays of S: pos_row (stage-2 row),
 * pos_col (Tbar column, -1 = RHS), kind (0 DISCRETE, 1 NORMAL, 2 UNIFORM), par_a / par_b (mean /
 * variance, left / right), cnt (outcomes of a DISCRETE element) and the rectangular
 * [S x MAX_OUTCOMES] tables of values and probabilities (zero padded).  Any pointer may be NULL. */
SQLP_API int32_t sqlp_smps_elements(sqlp_smps *smps, int32_t *pos_row, int32_t *pos_col,
                                    int32_t *kind, double *par_a, double *par_b, int32_t *cnt,
                                    double *vals, double *probs);
/* sdEpigraph(prob, w, lb) straight from the parsed files: sqlp_epi_create with the tables above,
 * then sqlp_epi_set_outcomes (cdf = left-to-right running sum of the probabilities) and, if any
 * element is continuous, sqlp_epi_set_distributions -- ready for sqlp_epi_sample_scenarios or
 * sqlp_epi_add_scenarios (values in element order). */
SQLP_API int32_t sqlp_epi_create_smps(sqlp_ctx *ctx, sqlp_pool *pool, sqlp_smps *smps,
                                      sqlp_epi **out);

/* eval_dual(coef, delta, x, dual) -- subprob.jl:128-131 -- for LOCAL scenario `scen` and
 * pool slot `vertex`, in the reference's operation order (debug / parity pin). */
SQLP_API int32_t sqlp_eval_dual(sqlp_epi *epi, int64_t local_scen, int64_t vertex,
                                const double *x, double *out);

#ifdef __cplusplus
}
#endif
#endif /* SQLP_B200_H */

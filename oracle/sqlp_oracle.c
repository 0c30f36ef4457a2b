/*
 * sqlp_oracle.c -- CPU ORACLE for the argmax cut-formation path of yhz0/SQLP (TwoSD).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (sqlp_b200/, libsqlp_b200.so) never links, imports or calls anything in oracle/.
 *
 * It restates, loop for loop, the reference's Julia code (paths relative to the
 * reference checkout):
 *   A1  delta_coefficients      src/sd_algorithm/subprob.jl:104-121
 *   A7  eval_dual               src/sd_algorithm/subprob.jl:128-131
 *   A2  argmax_procedure        src/sd_algorithm/subprob.jl:141-169
 *   A3  build_sasa_cut          src/sd_algorithm/epigraph.jl:125-146
 *   A5  hash_dual_vector/isequal/push!   src/sd_algorithm/dual_set.jl:24-53,84-93
 *
 * Third-party arithmetic the reference leans on, which is NOT under the reference
 * tree (Julia 1.9.3 Base/SparseArrays/LinearAlgebra, OpenBLAS 0.3.21 ddot) is
 * restated from its published algorithm:
 *   - Base.round(x; base=2, sigdigits=16): scale by the exact power of two that puts
 *     16 significant bits left of the binary point, round-to-nearest-even, scale back.
 *   - dot(::Vector,::Vector): plain index-order sum (OpenBLAS's SIMD order is
 *     architecture dependent, so bitwise parity with Julia is undefined even CPU to
 *     CPU; the parity rule therefore carries a 1e-12 relative argmax-gap exemption).
 *   - SparseMatrixCSC * Vector: column-ordered scatter y[row] += nz * x[col].
 *   - Adjoint(SparseMatrixCSC) * Vector: per column, row-ordered gather.
 *
 * Pinning: this file is checked against every known-answer vector the reference's own
 * tests hold for the path (test/dual_set_test.jl, test/sd_test.jl, test/sgd_example.jl;
 * see tests/test_oracle_golden.py).  Julia is not installed in this image, so beyond
 * those lands-sized vectors parity with the Julia runtime is UNPINNED; this restatement
 * is then the oracle of record.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -fPIC -shared, no -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* --------------------------------------------------------------------------------
 * A5: dedup rule, dual_set.jl
 * -------------------------------------------------------------------------------- */

/* dual_set.jl:4 */
#define ORC_SIGNIFICANT_DIGITS 16

static double orc_pow2(int e)
{
    /* 2.0^e as Julia's Float64^Int gives it: exact in range, Inf above, 0 below. */
    if (e > 1023) return INFINITY;
    if (e < -1074) return 0.0;
    return ldexp(1.0, e);
}

/* Base.round(x; base=2, sigdigits=16) -- call sites dual_set.jl:32-33,51. */
ORC_API double orc_round_sig(double x)
{
    if (!isfinite(x)) return x;
    int h = (x == 0.0) ? 0 : 1 + ilogb(x); /* hidigit(x, 2) = 1 + exponent(x) */
    int digits = ORC_SIGNIFICANT_DIGITS - h;
    double r;
    if (digits >= 0) {
        double sc = orc_pow2(digits);
        r = rint(x * sc) / sc;
    } else {
        double isc = orc_pow2(-digits);
        r = rint(x / isc) * isc;
    }
    if (!isfinite(r)) {
        if (digits > 0) return x;
        return (x > 0.0) ? 0.0 : ((x < 0.0) ? -0.0 : x);
    }
    return r;
}

/* hash_dual_vector, dual_set.jl:46-53: sequential 1-norm, rounded, reinterpreted. */
ORC_API uint64_t orc_hash_dual_vector(const double *v, int64_t n)
{
    double mysum = 0.0;
    for (int64_t i = 0; i < n; ++i) mysum += fabs(v[i]);
    double r = orc_round_sig(mysum);
    uint64_t h;
    memcpy(&h, &r, sizeof h);
    return h;
}

/* isequal, dual_set.jl:24-40. */
ORC_API int orc_isequal(const double *a, int64_t na, uint64_t ha,
                        const double *b, int64_t nb, uint64_t hb)
{
    if (na != nb || ha != hb) return 0;
    for (int64_t i = 0; i < na; ++i) {
        double r1 = orc_round_sig(a[i]);
        double r2 = orc_round_sig(b[i]);
        if (r1 != r2) return 0; /* NaN != NaN: a vector holding NaN never matches */
    }
    return 1;
}

/* push!, dual_set.jl:84-93 for a fixed-length pool stored row-major [cap][m2].
 * Returns the 0-based slot of v (new or the first duplicate); *inserted says which.
 * Returns -1 when the pool is full. */
ORC_API int64_t orc_pool_push(double *pool, uint64_t *hashes, int64_t *K, int64_t cap,
                              int64_t m2, const double *v, int32_t *inserted)
{
    uint64_t h = orc_hash_dual_vector(v, m2);
    for (int64_t k = 0; k < *K; ++k) {
        if (orc_isequal(v, m2, h, pool + k * m2, m2, hashes[k])) {
            *inserted = 0;
            return k;
        }
    }
    if (*K >= cap) return -1;
    memcpy(pool + (*K) * m2, v, (size_t)m2 * sizeof(double));
    hashes[*K] = h;
    *inserted = 1;
    return (*K)++;
}

/* --------------------------------------------------------------------------------
 * Problem description shared by A1/A2/A3/A7.
 *
 * coef   : rbar dense [m2] (non-stored entries of the reference's SparseVector are 0.0),
 *          Tbar CSC 0-based (colptr[n1+1], rowval, nzval), rows ascending per column.
 * table  : the stochastic-position table, s entries (pos_row[e], pos_col[e]); col -1 is
 *          the "RHS"/"rhs" column (subprob.jl:113).
 * values : realised values, row-major [N][s] in table order.
 * -------------------------------------------------------------------------------- */
typedef struct {
    int64_t m2, n1;
    const double *rbar;
    const int64_t *T_colptr;
    const int64_t *T_rowval;
    const double *T_nzval;
    int64_t s;
    const int32_t *pos_row;
    const int32_t *pos_col;
} orc_problem;

static double orc_T_entry(const orc_problem *P, int64_t row, int64_t col)
{
    for (int64_t k = P->T_colptr[col]; k < P->T_colptr[col + 1]; ++k)
        if (P->T_rowval[k] == row) return P->T_nzval[k];
    return 0.0;
}

/* A1 delta_coefficients, subprob.jl:104-121, for ONE scenario.
 * delta_rhs: dense [m2] (zero-filled here).  delta_T: one value per table entry
 * (0.0 for RHS entries), i.e. the (row, col, value) triplets of the reference's
 * sparse delta_transfer in table order. */
ORC_API void orc_delta_coefficients(const orc_problem *P, const double *values,
                                    double *delta_rhs, double *delta_T)
{
    memset(delta_rhs, 0, (size_t)P->m2 * sizeof(double));
    for (int64_t e = 0; e < P->s; ++e) {
        int64_t row = P->pos_row[e];
        if (P->pos_col[e] < 0) {
            delta_rhs[row] = values[e] - P->rbar[row]; /* :114 */
            delta_T[e] = 0.0;
        } else {
            delta_T[e] = values[e] - orc_T_entry(P, row, P->pos_col[e]); /* :117 */
        }
    }
}

/* order of the table's transfer entries inside a CSC matrix: by (col, row). */
typedef struct { int64_t col, row, e; } orc_trip;
static int orc_trip_cmp(const void *a, const void *b)
{
    const orc_trip *x = a, *y = b;
    if (x->col != y->col) return x->col < y->col ? -1 : 1;
    if (x->row != y->row) return x->row < y->row ? -1 : 1;
    return x->e < y->e ? -1 : (x->e > y->e);
}
static int64_t orc_transfer_order(const orc_problem *P, orc_trip **out)
{
    int64_t nt = 0;
    for (int64_t e = 0; e < P->s; ++e) nt += (P->pos_col[e] >= 0);
    orc_trip *t = malloc((size_t)(nt ? nt : 1) * sizeof *t);
    int64_t c = 0;
    for (int64_t e = 0; e < P->s; ++e)
        if (P->pos_col[e] >= 0) t[c++] = (orc_trip){P->pos_col[e], P->pos_row[e], e};
    qsort(t, (size_t)nt, sizeof *t, orc_trip_cmp);
    *out = t;
    return nt;
}

/* coef.transfer * x as SparseArrays does it: column-ordered scatter. */
static void orc_spmv_csc(const orc_problem *P, const double *x, double *y)
{
    memset(y, 0, (size_t)P->m2 * sizeof(double));
    for (int64_t j = 0; j < P->n1; ++j) {
        double xj = x[j];
        for (int64_t k = P->T_colptr[j]; k < P->T_colptr[j + 1]; ++k)
            y[P->T_rowval[k]] += P->T_nzval[k] * xj;
    }
}

/* dot(::Vector, ::Vector) restated as an index-order sum (oracle of record). */
static double orc_dot_seq(const double *a, const double *b, int64_t n)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* An 8-accumulator dot, the shape of an unrolled SIMD BLAS kernel.  Used ONLY by the
 * timing entry point (orc_bench_*) so the CPU baseline is not handicapped by the
 * latency-bound sequential sum; never used for parity. */
static double orc_dot_simd8(const double *a, const double *b, int64_t n)
{
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
    int64_t i = 0;
    for (; i + 8 <= n; i += 8) {
        s0 += a[i] * b[i];         s1 += a[i + 1] * b[i + 1];
        s2 += a[i + 2] * b[i + 2]; s3 += a[i + 3] * b[i + 3];
        s4 += a[i + 4] * b[i + 4]; s5 += a[i + 5] * b[i + 5];
        s6 += a[i + 6] * b[i + 6]; s7 += a[i + 7] * b[i + 7];
    }
    double s = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
    for (; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* delta.delta_rhs - delta.delta_transfer * x  (subprob.jl:149), dense [m2]. */
static void orc_delta_vector(const orc_problem *P, const orc_trip *ord, int64_t nt,
                             const double *delta_rhs, const double *delta_T,
                             const double *x, double *scratch, double *dvec)
{
    memset(scratch, 0, (size_t)P->m2 * sizeof(double));
    for (int64_t t = 0; t < nt; ++t)
        scratch[ord[t].row] += delta_T[ord[t].e] * x[ord[t].col];
    for (int64_t r = 0; r < P->m2; ++r) dvec[r] = delta_rhs[r] - scratch[r];
}

/* A7 eval_dual, subprob.jl:128-131:
 *   dot(dual, (rhs + delta_rhs) - (transfer + delta_transfer) * x)  */
ORC_API double orc_eval_dual(const orc_problem *P, const double *values,
                             const double *x, const double *dual)
{
    int64_t m2 = P->m2;
    double *drhs = malloc((size_t)m2 * sizeof(double));
    double *dT = malloc((size_t)(P->s ? P->s : 1) * sizeof(double));
    double *y = calloc((size_t)m2, sizeof(double));
    double *v = malloc((size_t)m2 * sizeof(double));
    orc_trip *ord;
    int64_t nt = orc_transfer_order(P, &ord);
    orc_delta_coefficients(P, values, drhs, dT);
    /* (transfer + delta_transfer) * x: the merged matrix (a cell present in both holds
     * T + dT, formed before the multiply), column-ordered scatter, rows ascending. */
    int64_t t = 0;
    for (int64_t j = 0; j < P->n1; ++j) {
        double xj = x[j];
        int64_t k = P->T_colptr[j], kend = P->T_colptr[j + 1];
        while (t < nt && ord[t].col < j) ++t;
        while (k < kend || (t < nt && ord[t].col == j)) {
            int64_t rk = (k < kend) ? P->T_rowval[k] : INT64_MAX;
            int64_t rt = (t < nt && ord[t].col == j) ? ord[t].row : INT64_MAX;
            if (rk < rt) {
                y[rk] += P->T_nzval[k] * xj; ++k;
            } else if (rt < rk) {
                y[rt] += dT[ord[t].e] * xj; ++t;
            } else {
                y[rk] += (P->T_nzval[k] + dT[ord[t].e]) * xj; ++k; ++t;
            }
        }
    }
    for (int64_t r = 0; r < m2; ++r) v[r] = (P->rbar[r] + drhs[r]) - y[r];
    double out = orc_dot_seq(dual, v, m2);
    free(drhs); free(dT); free(y); free(v); free(ord);
    return out;
}

/* A2 argmax_procedure, subprob.jl:141-169 (MIN_SENSE branch).
 * pool: row-major [K][m2], iteration = insertion order (dual_set.jl:116-122).
 * Outputs max_val[N], max_idx[N] (0-based; -1 where no vertex ever beat -Inf, the
 * reference's undefined Ref), and optionally second[N] = best score among the
 * vertices that were NOT selected (for the parity rule's gap exemption).
 * dot_kind: 0 = index-order dot (oracle of record), 1 = 8-accumulator dot (timing). */
/* forced (may be NULL): scenario i is scored against vertex forced[i] alone -- what a checker needs when
 * the selection is given (the cut on a validated selection), O(N m2) instead of O(N K m2). */
static void orc_argmax_range_f(const orc_problem *P, int64_t i0, int64_t i1,
                               const double *values, const double *x, const double *pool,
                               int64_t K, double *max_val, int64_t *max_idx,
                               double *second, int dot_kind, const int64_t *forced);

static void orc_argmax_range(const orc_problem *P, int64_t i0, int64_t i1,
                             const double *values, const double *x, const double *pool,
                             int64_t K, double *max_val, int64_t *max_idx,
                             double *second, int dot_kind)
{
    orc_argmax_range_f(P, i0, i1, values, x, pool, K, max_val, max_idx, second, dot_kind, NULL);
}

static void orc_argmax_range_f(const orc_problem *P, int64_t i0, int64_t i1,
                               const double *values, const double *x, const double *pool,
                               int64_t K, double *max_val, int64_t *max_idx,
                               double *second, int dot_kind, const int64_t *forced)
{
    int64_t m2 = P->m2;
    double *base = malloc((size_t)m2 * sizeof(double));
    double *tx = malloc((size_t)m2 * sizeof(double));
    double *drhs = malloc((size_t)m2 * sizeof(double));
    double *dT = malloc((size_t)(P->s ? P->s : 1) * sizeof(double));
    double *scratch = malloc((size_t)m2 * sizeof(double));
    double *dvec = malloc((size_t)m2 * sizeof(double));
    orc_trip *ord;
    int64_t nt = orc_transfer_order(P, &ord);

    orc_spmv_csc(P, x, tx);                                     /* :147 */
    for (int64_t r = 0; r < m2; ++r) base[r] = P->rbar[r] - tx[r];

    for (int64_t i = i0; i < i1; ++i) {                         /* :148 */
        orc_delta_coefficients(P, values + i * P->s, drhs, dT); /* stored at add_scenario! */
        orc_delta_vector(P, ord, nt, drhs, dT, x, scratch, dvec); /* :149 */
        double cur = -INFINITY, sec = -INFINITY;                /* :151 */
        int64_t arg = -1;
        const int64_t kb = forced ? forced[i] : 0, ke = forced ? forced[i] + 1 : K;
        for (int64_t k = (kb < 0 ? 0 : kb); k < (kb < 0 ? 0 : ke); ++k) {   /* :154 */
            const double *p = pool + k * m2;
            double v = dot_kind ? orc_dot_simd8(p, base, m2) + orc_dot_simd8(p, dvec, m2)
                                : orc_dot_seq(p, base, m2) + orc_dot_seq(p, dvec, m2); /* :155 */
            if (v > cur) {                                      /* :156 strict: first max wins */
                sec = cur;
                cur = v;
                arg = k;
            } else if (v > sec) {
                sec = v;
            }
        }
        max_val[i] = cur;
        max_idx[i] = arg;
        if (second) second[i] = sec;
    }
    free(base); free(tx); free(drhs); free(dT); free(scratch); free(dvec); free(ord);
}

ORC_API void orc_argmax_procedure(const orc_problem *P, int64_t N, const double *values,
                                  const double *x, const double *pool, int64_t K,
                                  double *max_val, int64_t *max_idx, double *second)
{
    orc_argmax_range(P, 0, N, values, x, pool, K, max_val, max_idx, second, 0);
}

/* Score of one (scenario, vertex) pair exactly as A2 forms it (two dots), plus a
 * long-double evaluation of the same quantity for tolerance calibration. */
ORC_API void orc_score_pair(const orc_problem *P, const double *values, const double *x,
                            const double *vertex, double *score, long double *score_ld)
{
    int64_t m2 = P->m2;
    double *base = malloc((size_t)m2 * sizeof(double));
    double *tx = malloc((size_t)m2 * sizeof(double));
    double *drhs = malloc((size_t)m2 * sizeof(double));
    double *dT = malloc((size_t)(P->s ? P->s : 1) * sizeof(double));
    double *scratch = malloc((size_t)m2 * sizeof(double));
    double *dvec = malloc((size_t)m2 * sizeof(double));
    orc_trip *ord;
    int64_t nt = orc_transfer_order(P, &ord);
    orc_spmv_csc(P, x, tx);
    for (int64_t r = 0; r < m2; ++r) base[r] = P->rbar[r] - tx[r];
    orc_delta_coefficients(P, values, drhs, dT);
    orc_delta_vector(P, ord, nt, drhs, dT, x, scratch, dvec);
    if (score) *score = orc_dot_seq(vertex, base, m2) + orc_dot_seq(vertex, dvec, m2);
    if (score_ld) {
        long double acc = 0.0L;
        for (int64_t r = 0; r < m2; ++r)
            acc += (long double)vertex[r] * ((long double)base[r] + (long double)dvec[r]);
        *score_ld = acc;
    }
    free(base); free(tx); free(drhs); free(dT); free(scratch); free(dvec); free(ord);
}

/* A3 build_sasa_cut, epigraph.jl:125-146.
 * weights[N] = epi.scenario_weight, total_weight = epi.total_scenario_weight.
 * Outputs alpha, beta[n1], weight_mark, val (= sum p_i maxval_i, the dead variable of
 * :142, kept because alpha + beta.x == val is invariant G5), and optionally the
 * argmax results.  Returns 0, or -1 if some scenario had no argmax (UndefRefError in
 * the reference at :140).
 * forced_idx (normally NULL): aggregate THESE vertices instead of the oracle's own argmax.
 * Used by the parity tests after the device's selection passed the north-star rule
 * (identical, or within the 1e-12 relative score gap): on exact or near ties two valid
 * selections give different (alpha, beta) with the same alpha + beta.x, so coefficients
 * are compared on the same selection. */
ORC_API int32_t orc_build_sasa_cut(const orc_problem *P, int64_t N, const double *values,
                                   const double *weights, double total_weight,
                                   const double *x, const double *pool, int64_t K,
                                   double *alpha_out, double *beta_out,
                                   double *weight_mark, double *val_out,
                                   double *max_val_out, int64_t *max_idx_out,
                                   const int64_t *forced_idx)
{
    int64_t m2 = P->m2, n1 = P->n1;
    double *max_val = max_val_out ? max_val_out : malloc((size_t)(N ? N : 1) * sizeof(double));
    int64_t *max_idx = max_idx_out ? max_idx_out : malloc((size_t)(N ? N : 1) * sizeof(int64_t));
    /* :127 -- with a forced selection only the selected vertices are scored (max_val = their scores) */
    orc_argmax_range_f(P, 0, N, values, x, pool, K, max_val, max_idx, NULL, 0, forced_idx);

    double alpha = 0.0, val = 0.0;                               /* :130-132 */
    for (int64_t j = 0; j < n1; ++j) beta_out[j] = 0.0;
    double *drhs = malloc((size_t)m2 * sizeof(double));
    double *dT = malloc((size_t)(P->s ? P->s : 1) * sizeof(double));
    orc_trip *ord;
    int64_t nt = orc_transfer_order(P, &ord);
    int32_t status = 0;

    for (int64_t i = 0; i < N; ++i) {                            /* :134 */
        int64_t sel = forced_idx ? forced_idx[i] : max_idx[i];
        if (sel < 0) { status = -1; break; }
        const double *dual = pool + sel * m2;                    /* :136 */
        double p = weights[i] / total_weight;                    /* :138 */
        orc_delta_coefficients(P, values + i * P->s, drhs, dT);

        /* :140  alpha += p * dot(dual, rhs + delta_rhs)   (index order over the union
         * of stored entries; absent entries contribute dual*0 = 0) */
        double d = 0.0;
        for (int64_t r = 0; r < m2; ++r) {
            double rr = P->rbar[r] + drhs[r];
            if (rr != 0.0) d += dual[r] * rr;
        }
        alpha += p * d;

        /* :141  beta += -p * (transfer + delta_transfer)' * dual
         * per column: row-ordered gather over the merged pattern, then scale by -p. */
        int64_t t = 0;
        for (int64_t j = 0; j < n1; ++j) {
            double tmp = 0.0;
            int64_t k = P->T_colptr[j], kend = P->T_colptr[j + 1];
            while (t < nt && ord[t].col < j) ++t;
            int64_t tt = t;
            while (k < kend || (tt < nt && ord[tt].col == j)) {
                int64_t rk = (k < kend) ? P->T_rowval[k] : INT64_MAX;
                int64_t rt = (tt < nt && ord[tt].col == j) ? ord[tt].row : INT64_MAX;
                if (rk < rt) {
                    tmp += P->T_nzval[k] * dual[rk]; ++k;
                } else if (rt < rk) {
                    tmp += dT[ord[tt].e] * dual[rt]; ++tt;
                } else {
                    tmp += (P->T_nzval[k] + dT[ord[tt].e]) * dual[rk]; ++k; ++tt;
                }
            }
            beta_out[j] += tmp * (-p);
        }
        val += p * max_val[i];                                   /* :142 */
    }
    *alpha_out = alpha;
    *weight_mark = total_weight;                                 /* :145 */
    if (val_out) *val_out = val;
    free(drhs); free(dT); free(ord);
    if (!max_val_out) free(max_val);
    if (!max_idx_out) free(max_idx);
    return status;
}

/* --------------------------------------------------------------------------------
 * Timing entry point for bench.py's cpu_baseline / --impl reference legs.
 * Same loop structure as A2 (scenario-outer, vertex-inner, dense length-m2 base and
 * delta vectors, two dots per pair, strict > first-max), scenarios split over OpenMP
 * threads (the reference is single threaded; threads>1 is a courtesy upper bound for a
 * threaded port).  dot_kind as above.  Returns the number of threads used.
 * -------------------------------------------------------------------------------- */
ORC_API int32_t orc_bench_argmax(const orc_problem *P, int64_t N, const double *values,
                                 const double *x, const double *pool, int64_t K,
                                 double *max_val, int64_t *max_idx, int32_t threads,
                                 int32_t dot_kind)
{
    int used = 1;
#ifdef _OPENMP
    if (threads < 1) threads = omp_get_max_threads();
    used = threads;
#pragma omp parallel num_threads(threads)
    {
        int nt = omp_get_num_threads(), id = omp_get_thread_num();
        int64_t per = (N + nt - 1) / nt;
        int64_t i0 = id * per, i1 = i0 + per;
        if (i1 > N) i1 = N;
        if (i0 < i1)
            orc_argmax_range(P, i0, i1, values, x, pool, K, max_val, max_idx, NULL, dot_kind);
    }
#else
    (void)threads;
    orc_argmax_range(P, 0, N, values, x, pool, K, max_val, max_idx, NULL, dot_kind);
#endif
    return used;
}

ORC_API int32_t orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* --------------------------------------------------------------------------------
 * Counter-based generator shared with the device:
 *   u(seed, idx) = top 53 bits of splitmix64(seed ^ idx*0x9E3779B97F4A7C15) / 2^53
 * -------------------------------------------------------------------------------- */
ORC_API double orc_u01(uint64_t seed, uint64_t idx)
{
    uint64_t z = seed ^ (idx * 0x9E3779B97F4A7C15ULL);
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

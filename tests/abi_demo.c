/*
 * abi_demo.c -- the C ABI of libsqlp_b200.so driven from plain C (no Python, no torch): the calls a
 * non-Julia host would make for the reference's build_sasa_cut test (test/sd_test.jl:207-235) --
 * two dual vertices, two weighted scenarios of lands (RHS of row S2C5 = 3 and 7, weights 1.5 and
 * 0.5), one cut at x = [2, 3, 4, 5].  Prints one JSON object; tests/test_gpu_parity.py compiles
 * it with gcc, runs it and checks the reference's closed form.
 */
#include <stdio.h>
#include <stdlib.h>

#include "../include/sqlp_b200.h"

#define OK(call)                                                                        \
    do {                                                                                \
        int32_t rc_ = (call);                                                           \
        if (rc_ != SQLP_OK) {                                                           \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)rc_, sqlp_last_error());      \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

int main(void)
{
    /* lands, second stage: m2 = 7 rows, n1 = 4 first-stage columns (tests/golden/instances/lands.npz) */
    const int64_t m2 = 7, n1 = 4;
    const int64_t r_idx[2] = {5, 6};
    const double r_val[2] = {3.0, 2.0};
    const int64_t T_colptr[5] = {0, 1, 2, 3, 4}, T_rowval[4] = {0, 1, 2, 3};
    const double T_nzval[4] = {-1.0, -1.0, -1.0, -1.0};
    const int32_t pos_row[1] = {4}, pos_col[1] = {-1}; /* the random element: RHS of S2C5 */
    const double my_dual[7] = {-4.0, -1.0, -12.0, -0.0, 44.0, 28.0, 5.5};
    const double my_dual_2[7] = {-0.5, 0.0, -8.5, 0.0, 40.5, 24.5, 4.5};
    const double values[2] = {3.0, 7.0}, weights[2] = {1.5, 0.5}, x[4] = {2.0, 3.0, 4.0, 5.0};

    sqlp_ctx *ctx = NULL;
    sqlp_pool *pool = NULL;
    sqlp_epi *epi = NULL;
    OK(sqlp_ctx_create(0, &ctx));
    OK(sqlp_pool_create(ctx, m2, &pool));
    int32_t ins[3];
    int64_t idx[3];
    OK(sqlp_pool_push(pool, my_dual, &ins[0], &idx[0]));
    OK(sqlp_pool_push(pool, my_dual_2, &ins[1], &idx[1]));
    OK(sqlp_pool_push(pool, my_dual, &ins[2], &idx[2])); /* a duplicate */
    int64_t K = 0;
    OK(sqlp_pool_size(pool, &K));
    OK(sqlp_epi_create(ctx, pool, m2, n1, 2, r_idx, r_val, T_colptr, T_rowval, T_nzval, 1, pos_row, pos_col, &epi));
    OK(sqlp_epi_add_scenarios(epi, 2, values, weights));
    double max_val[2], alpha = 0, beta[4], wm = 0, val = 0;
    int64_t max_idx[2];
    OK(sqlp_epi_argmax(epi, x, SQLP_MIN_SENSE, max_val, max_idx));
    OK(sqlp_epi_build_cut(epi, x, &alpha, beta, &wm, &val));
    int32_t unsupported = sqlp_epi_argmax(epi, x, SQLP_MAX_SENSE, max_val, max_idx);
    printf("{\"version\": \"%s\", \"K\": %lld, \"inserted\": [%d, %d, %d], \"index\": [%lld, %lld, %lld], "
           "\"max_val\": [%.17g, %.17g], \"max_idx\": [%lld, %lld], \"alpha\": %.17g, "
           "\"beta\": [%.17g, %.17g, %.17g, %.17g], \"weight_mark\": %.17g, \"val\": %.17g, \"max_sense_status\": %d}\n",
           sqlp_version(), (long long)K, ins[0], ins[1], ins[2], (long long)idx[0], (long long)idx[1],
           (long long)idx[2], max_val[0], max_val[1], (long long)max_idx[0], (long long)max_idx[1], alpha, beta[0],
           beta[1], beta[2], beta[3], wm, val, (int)unsupported);
    OK(sqlp_epi_destroy(epi));
    OK(sqlp_pool_destroy(pool));
    OK(sqlp_ctx_destroy(ctx));
    return 0;
}

/* The SMPS reader of libsqlp_b200.so driven from plain C against include/sqlp_b200.h only (row N4 of the
 * scope table; needs no GPU): load cor / tim / sto, print the dimensions, the stage-2 tables and the random
 * elements as one JSON object.  usage: abi_smps_demo file.cor file.tim file.sto */
#include <stdio.h>
#include <stdlib.h>
#include "../include/sqlp_b200.h"

#define CHECK(call)                                                            \
    do {                                                                       \
        int32_t st_ = (call);                                                  \
        if (st_ != SQLP_OK) {                                                  \
            fprintf(stderr, "%s -> %d: %s\n", #call, st_, sqlp_last_error()); \
            return 1;                                                          \
        }                                                                      \
    } while (0)

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    sqlp_smps *p = NULL;
    CHECK(sqlp_smps_load(argv[1], argv[2], argv[3], &p));
    int64_t d[SQLP_SMPS_NDIMS];
    CHECK(sqlp_smps_dims(p, d));
    const int64_t n1 = d[SQLP_SMPS_N1], m2 = d[SQLP_SMPS_M2], s = d[SQLP_SMPS_S], mo = d[SQLP_SMPS_MAX_OUTCOMES];
    double *rbar = malloc(sizeof(double) * (size_t)(m2 + 1));
    int64_t *colptr = malloc(sizeof(int64_t) * (size_t)(n1 + 1));
    int64_t *rowval = malloc(sizeof(int64_t) * (size_t)(d[SQLP_SMPS_T_NNZ] + 1));
    double *nzval = malloc(sizeof(double) * (size_t)(d[SQLP_SMPS_T_NNZ] + 1));
    CHECK(sqlp_smps_stage2(p, rbar, colptr, rowval, nzval, NULL, NULL, NULL, NULL, NULL));
    int32_t *pos_row = malloc(sizeof(int32_t) * (size_t)(s + 1)), *pos_col = malloc(sizeof(int32_t) * (size_t)(s + 1));
    int32_t *kind = malloc(sizeof(int32_t) * (size_t)(s + 1)), *cnt = malloc(sizeof(int32_t) * (size_t)(s + 1));
    double *vals = malloc(sizeof(double) * (size_t)(s * mo + 1)), *probs = malloc(sizeof(double) * (size_t)(s * mo + 1));
    CHECK(sqlp_smps_elements(p, pos_row, pos_col, kind, NULL, NULL, cnt, vals, probs));
    char name[128], row[128], col[128];
    CHECK(sqlp_smps_name(p, SQLP_SMPS_COR_NAME, 0, name, sizeof name));
    printf("{\"name\": \"%s\", \"n1\": %lld, \"n2\": %lld, \"m2\": %lld, \"s\": %lld, \"T_nnz\": %lld, \"rbar\": [", name,
           (long long)n1, (long long)d[SQLP_SMPS_N2], (long long)m2, (long long)s, (long long)d[SQLP_SMPS_T_NNZ]);
    for (int64_t i = 0; i < m2; ++i) printf("%s%.17g", i ? ", " : "", rbar[i]);
    printf("], \"T\": [");
    for (int64_t j = 0, first = 1; j < n1; ++j)
        for (int64_t q = colptr[j]; q < colptr[j + 1]; ++q, first = 0)
            printf("%s[%lld, %lld, %.17g]", first ? "" : ", ", (long long)rowval[q], (long long)j, nzval[q]);
    printf("], \"elements\": [");
    for (int64_t e = 0; e < s; ++e) {
        CHECK(sqlp_smps_name(p, SQLP_SMPS_ELEM_COL, e, col, sizeof col));
        CHECK(sqlp_smps_name(p, SQLP_SMPS_ELEM_ROW, e, row, sizeof row));
        printf("%s{\"col\": \"%s\", \"row\": \"%s\", \"pos\": [%d, %d], \"kind\": %d, \"outcomes\": [", e ? ", " : "", col, row,
               pos_row[e], pos_col[e], kind[e]);
        for (int32_t o = 0; o < cnt[e]; ++o)
            printf("%s[%.17g, %.17g]", o ? ", " : "", vals[e * mo + o], probs[e * mo + o]);
        printf("]}");
    }
    printf("]}\n");
    /* a reader error comes back as a status and a message, not as an exception */
    sqlp_smps *bad = NULL;
    if (sqlp_smps_load("/nonexistent.cor", argv[2], argv[3], &bad) != SQLP_E_IO || bad) return 3;
    CHECK(sqlp_smps_destroy(p));
    return 0;
}

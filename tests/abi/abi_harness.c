/*
 * abi_harness.c — drives libsevero_b200.so the way Julia's `ccall` would, with no Python in the process: plain C, the public
 * header only, Julia-style arrays (1-based Int64 colptr / rowval, Float64 nzval, column-major dense outputs). It replaces the
 * call of src/irlba.jl:66-71: upload a SparseMatrixCSC, wrap it as CenteredMatrix (scaling.jl:219-232), run `irlba`, and
 * check the result with the reference's own criterion (test/test_irlba.jl:30) in plain loops.
 *
 *   gcc -O2 -I include tests/abi/abi_harness.c -o abi_harness -L severo.jl_b200 -lsevero_b200 -Wl,-rpath,$PWD/severo.jl_b200 -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "severo_b200.h"

#define CHECK(call)                                                                 \
    do {                                                                            \
        int rc_ = (call);                                                           \
        if (rc_ != SVB_OK) {                                                        \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, svb_last_error());        \
            return 1;                                                               \
        }                                                                           \
    } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double urand(void) {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) * (1.0 / 9007199254740992.0);
}

int main(void) {
    const int64_t m = 4000, n = 300, nu = 6;
    /* a random sparse matrix with a planted rank-nu part, built column by column (CSC, rows ascending) */
    int64_t *colptr = malloc((size_t)(n + 1) * sizeof(int64_t));
    int64_t cap = m * n / 8, nnz = 0;
    int64_t *rowval = malloc((size_t)cap * sizeof(int64_t));
    double *nzval = malloc((size_t)cap * sizeof(double));
    double *a = malloc((size_t)m * nu * sizeof(double)), *b = malloc((size_t)n * nu * sizeof(double));
    for (int64_t i = 0; i < m * nu; ++i) a[i] = urand() - 0.5;
    for (int64_t i = 0; i < n * nu; ++i) b[i] = urand() - 0.5;
    for (int64_t j = 0; j < n; ++j) {
        colptr[j] = nnz + 1; /* 1-based, as Julia holds it */
        for (int64_t i = 0; i < m && nnz < cap; ++i) {
            if (urand() < 0.06) {
                double v = 0.1 * (urand() - 0.5);
                for (int64_t k = 0; k < nu; ++k) v += (double)(nu - k) * a[i + k * m] * b[j + k * n];
                rowval[nnz] = i + 1;
                nzval[nnz] = v;
                ++nnz;
            }
        }
    }
    colptr[n] = nnz + 1;
    double *mu = malloc((size_t)n * sizeof(double)); /* column means: S = A - 1*mu' is the column-centred matrix */
    for (int64_t j = 0; j < n; ++j) {
        double s = 0.0;
        for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) s += nzval[p];
        mu[j] = s / (double)m;
    }

    CHECK(svb_init(0));
    svb_matrix_t A = NULL;
    CHECK(svb_csc_upload(m, n, colptr, rowval, SVB_I64, nzval, SVB_F64, 1, &A));
    int64_t r_ = 0, c_ = 0, z_ = 0;
    int vt = 0;
    CHECK(svb_matrix_info(A, &r_, &c_, &z_, &vt));
    if (r_ != m || c_ != n || z_ != nnz || vt != SVB_F64) { fprintf(stderr, "matrix_info mismatch\n"); return 1; }
    svb_operator_t S = NULL;
    CHECK(svb_operator_create(A, mu, 0, &S));
    CHECK(svb_matrix_free(A));

    double *init = malloc((size_t)n * sizeof(double));
    for (int64_t j = 0; j < n; ++j) init[j] = urand() - 0.5;
    double *s = calloc((size_t)nu, sizeof(double)), *U = calloc((size_t)m * nu, sizeof(double)), *V = calloc((size_t)n * nu, sizeof(double));
    int64_t iter = 0, mprod = 0;
    const double tol = 1e-9;
    CHECK(svb_irlba(S, nu, nu + 7, 1000, 0, tol, tol, init, s, U, V, &iter, &mprod));

    /* test/test_irlba.jl:30: ||S'U - V Sigma|| / ||S|| < tol, with S applied by plain loops on the host copy */
    double fro2 = 0.0, res2 = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        double colsq = 0.0, colsum = 0.0;
        for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) { colsq += nzval[p] * nzval[p]; colsum += nzval[p]; }
        fro2 += colsq - 2.0 * mu[j] * colsum + (double)m * mu[j] * mu[j];
        for (int64_t k = 0; k < nu; ++k) {
            double acc = 0.0, usum = 0.0;
            for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; ++p) acc += nzval[p] * U[(rowval[p] - 1) + k * m];
            for (int64_t i = 0; i < m; ++i) usum += U[i + k * m];
            const double d = (acc - mu[j] * usum) - V[j + k * n] * s[k];
            res2 += d * d;
        }
    }
    const double rel = sqrt(res2 / fro2);
    double orth = 0.0;
    for (int64_t k = 0; k < nu; ++k)
        for (int64_t l = 0; l < nu; ++l) {
            double d = 0.0;
            for (int64_t j = 0; j < n; ++j) d += V[j + k * n] * V[j + l * n];
            d -= (k == l) ? 1.0 : 0.0;
            if (fabs(d) > orth) orth = fabs(d);
        }
    int sorted = 1;
    for (int64_t k = 1; k < nu; ++k) sorted &= (s[k] <= s[k - 1]) && s[k] > 0.0;
    printf("irlba through the C ABI: sigma_1 = %.12g sigma_nu = %.12g, %lld restarts, %lld mat-vecs, residual %.3e, orth %.3e\n", s[0],
           s[nu - 1], (long long)iter, (long long)mprod, rel, orth);
    CHECK(svb_operator_free(S));
    CHECK(svb_shutdown());
    if (!(rel < tol) || !(orth < 1e-10) || !sorted) { fprintf(stderr, "ABI HARNESS FAILED\n"); return 1; }
    /* error convention: a bad argument returns a negative code and a message, never aborts */
    if (svb_init(0) != SVB_OK) return 1;
    svb_matrix_t bad = NULL;
    colptr[0] = 5; /* not a valid first pointer */
    const int rc = svb_csc_upload(m, n, colptr, rowval, SVB_I64, nzval, SVB_F64, 1, &bad);
    if (rc == SVB_OK || svb_last_error()[0] == 0) { fprintf(stderr, "malformed colptr accepted\n"); return 1; }
    svb_shutdown();
    printf("ABI HARNESS OK\n");
    return 0;
}

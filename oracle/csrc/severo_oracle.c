/*
 * severo_oracle.c — CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the Float64 loops of ExaScience/Severo.jl's PCA hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library. The product (severo.jl_b200) never links or calls it.
 *
 * Compile with -ffp-contract=off: Julia does not contract a*b+c into an FMA, GCC does by
 * default on x86-64 with -march=native; contraction would break bit-parity of the
 * order-exact loops (Welford, sf*x/s, x/std).
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference repo root). Index arrays are 0-based int64 here; the Python wrapper converts.
 *
 * Parity status: the pre-processing loops are pinned by the reference's fixed-matrix tests
 * (test/test_scaling.jl:22-45, test/test_input.jl:47-75). The IRLBA iterates live in the
 * external, un-vendored, un-pinned `libcell` (Severo_jll) => "parity unpinned" for
 * iterates; only converged results are pinned (test/test_irlba.jl vs dense SVD).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

EXPORT int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

EXPORT void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- normalize.jl:24  s = sum(A, dims=2) — exact Int64 row sums of the count matrix ---- */
EXPORT void orc_row_sums_i64(int64_t ncol, const int64_t *colptr, const int64_t *rowval,
                             const int64_t *nzval, int64_t nrow, int64_t *s) {
    memset(s, 0, (size_t)nrow * sizeof(int64_t));
    for (int64_t c = 0; c < ncol; ++c)
        for (int64_t j = colptr[c]; j < colptr[c + 1]; ++j) s[rowval[j]] += nzval[j];
}

/* ---- normalize.jl:25-29  nzB[j] = scale_factor * nzA[j] / s[rv[j]]  (one rounded multiply,
 *      then one rounded divide), and normalize.jl:36 log1p in place when do_log != 0 ---- */
EXPORT void orc_row_norm_f64(int64_t nnz, const int64_t *rowval, const int64_t *nzval,
                             const int64_t *s, double scale_factor, int do_log, double *out) {
    for (int64_t j = 0; j < nnz; ++j) {
        double t = scale_factor * (double)nzval[j];
        double v = t / (double)s[rowval[j]];
        out[j] = do_log ? log1p(v) : v;
    }
}

EXPORT void orc_row_norm_f32(int64_t nnz, const int64_t *rowval, const int64_t *nzval,
                             const int64_t *s, float scale_factor, int do_log, float *out) {
    for (int64_t j = 0; j < nnz; ++j) {
        float t = scale_factor * (float)nzval[j];
        float v = t / (float)s[rowval[j]];
        out[j] = do_log ? log1pf(v) : v;
    }
}

/* float-valued input variant (row_norm on an already-float matrix: sum is a float sum in
 * storage order per row as Julia's sum(A,dims=2) accumulates column by column) */
EXPORT void orc_row_sums_f64(int64_t ncol, const int64_t *colptr, const int64_t *rowval,
                             const double *nzval, int64_t nrow, double *s) {
    memset(s, 0, (size_t)nrow * sizeof(double));
    for (int64_t c = 0; c < ncol; ++c)
        for (int64_t j = colptr[c]; j < colptr[c + 1]; ++j) s[rowval[j]] += nzval[j];
}

/* ---- scaling.jl:18-34  mean_var(T, x): sequential Welford over the stored values of one
 *      column, `count` pre-seeded with the number of implicit zeros (scaling.jl:21) ---- */
static inline void welford_f64(const double *v, int64_t nnz, int64_t n, double *mu_out,
                               double *var_out) {
    int64_t count = n - nnz;
    double mu = 0.0, s = 0.0;
    for (int64_t k = 0; k < nnz; ++k) {
        count += 1;
        double delta = v[k] - mu;
        mu += delta / (double)count;
        s += delta * (v[k] - mu);
    }
    *mu_out = mu;
    *var_out = s / (double)(n - 1);
}

static inline void welford_i64(const int64_t *v, int64_t nnz, int64_t n, double *mu_out,
                               double *var_out) {
    int64_t count = n - nnz;
    double mu = 0.0, s = 0.0;
    for (int64_t k = 0; k < nnz; ++k) {
        count += 1;
        double x = (double)v[k];
        double delta = x - mu;
        mu += delta / (double)count;
        s += delta * (x - mu);
    }
    *mu_out = mu;
    *var_out = s / (double)(n - 1);
}

static inline void welford_f32(const float *v, int64_t nnz, int64_t n, float *mu_out,
                               float *var_out) {
    int64_t count = n - nnz;
    float mu = 0.0f, s = 0.0f;
    for (int64_t k = 0; k < nnz; ++k) {
        count += 1;
        float delta = v[k] - mu;
        mu += delta / (float)count;
        s += delta * (v[k] - mu);
    }
    *mu_out = mu;
    *var_out = s / (float)(n - 1);
}

/* scaling.jl:132-142 mean_var(T, A::SparseMatrixCSC): per column. Columns are independent, so
 * an OpenMP loop over columns does not change any bit. */
EXPORT void orc_mean_var_csc_f64(int64_t nrow, int64_t ncol, const int64_t *colptr,
                                 const double *nzval, double *mu, double *var) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t c = 0; c < ncol; ++c)
        welford_f64(nzval + colptr[c], colptr[c + 1] - colptr[c], nrow, mu + c, var + c);
}

EXPORT void orc_mean_var_csc_i64(int64_t nrow, int64_t ncol, const int64_t *colptr,
                                 const int64_t *nzval, double *mu, double *var) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t c = 0; c < ncol; ++c)
        welford_i64(nzval + colptr[c], colptr[c + 1] - colptr[c], nrow, mu + c, var + c);
}

EXPORT void orc_mean_var_csc_f32(int64_t nrow, int64_t ncol, const int64_t *colptr,
                                 const float *nzval, float *mu, float *var) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t c = 0; c < ncol; ++c)
        welford_f32(nzval + colptr[c], colptr[c + 1] - colptr[c], nrow, mu + c, var + c);
}

/* ---- variablefeatures.jl:19-28 standardized_var_clipped. Julia's `sum` is pairwise with an
 *      unspecified SIMD order; here the per-column sum is accumulated in long double and
 *      rounded once, i.e. this returns (to <=1 ulp) the correctly-rounded value that every
 *      summation order must be close to. ---- */
EXPORT void orc_stdvar_clipped_i64(int64_t nrow, int64_t ncol, const int64_t *colptr,
                                   const int64_t *nzval, const double *mu, const double *sd,
                                   double vmax, double *out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t c = 0; c < ncol; ++c) {
        out[c] = 0.0;
        if (sd[c] == 0.0) continue; /* variablefeatures.jl:24 */
        int64_t nnz = colptr[c + 1] - colptr[c];
        long double acc = 0.0L;
        for (int64_t j = colptr[c]; j < colptr[c + 1]; ++j) {
            double z = ((double)nzval[j] - mu[c]) / sd[c];
            if (z > vmax) z = vmax; /* standardize_clip = min(., vmax) variablefeatures.jl:19 */
            double z2 = z * z;
            acc += (long double)z2;
        }
        double z0 = (0.0 - mu[c]) / sd[c];
        if (z0 > vmax) z0 = vmax;
        double zterm = (double)(nrow - nnz) * (z0 * z0);
        double total = (double)acc + zterm;
        out[c] = total / (double)(nrow - 1);
    }
}

/* ---- scaling.jl:199-217 scale_data: per column Welford mean/std; stored mu = mean/std;
 *      B = min(x/std, scale_max + mu) on the stored entries only ---- */
EXPORT void orc_scale_data_f64(int64_t nrow, int64_t ncol, const int64_t *colptr,
                               const double *nzval, double scale_max, double *out,
                               double *mu_out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t c = 0; c < ncol; ++c) {
        double mu, var;
        welford_f64(nzval + colptr[c], colptr[c + 1] - colptr[c], nrow, &mu, &var);
        double sd = sqrt(var);
        mu = mu / sd;
        mu_out[c] = mu;
        double smax = scale_max + mu;
        for (int64_t j = colptr[c]; j < colptr[c + 1]; ++j) {
            double v = nzval[j] / sd;
            out[j] = (v > smax) ? smax : v;
        }
    }
}

/* integer input: mean_std(view(A,:,i)) runs in Float64 (scaling.jl:41), output dtype R */
EXPORT void orc_scale_data_i64(int64_t nrow, int64_t ncol, const int64_t *colptr,
                               const int64_t *nzval, double scale_max, double *out,
                               double *mu_out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t c = 0; c < ncol; ++c) {
        double mu, var;
        welford_i64(nzval + colptr[c], colptr[c + 1] - colptr[c], nrow, &mu, &var);
        double sd = sqrt(var);
        mu = mu / sd;
        mu_out[c] = mu;
        double smax = scale_max + mu;
        for (int64_t j = colptr[c]; j < colptr[c + 1]; ++j) {
            double v = (double)nzval[j] / sd;
            out[j] = (v > smax) ? smax : v;
        }
    }
}

/* ---- Julia stdlib SparseArrays mul!(C, A::CSC, v, a, b) as called from scaling.jl:247:
 *      C = b*C; for col: axj = v[col]*a; for j in col: C[rv[j]] += nzv[j]*axj.
 *      Serial scatter, ascending column order per row: exactly the stdlib order. ---- */
EXPORT void orc_csc_mul(int64_t nrow, int64_t ncol, const int64_t *colptr,
                        const int64_t *rowval, const double *nzval, const double *v, double alpha,
                        double beta, double *y) {
    if (beta == 0.0)
        memset(y, 0, (size_t)nrow * sizeof(double));
    else if (beta != 1.0)
        for (int64_t i = 0; i < nrow; ++i) y[i] *= beta;
    for (int64_t c = 0; c < ncol; ++c) {
        double axj = v[c] * alpha;
        for (int64_t j = colptr[c]; j < colptr[c + 1]; ++j) y[rowval[j]] += nzval[j] * axj;
    }
}

/* Same product from a CSR-by-row copy: y_i = sum_j (ascending j) a_ij * (v_j*alpha). With
 * beta == 0 the additions per row happen in the same order as the scatter above, so the
 * result is bit-identical; this form parallelises over rows (used for the all-cores CPU
 * baseline; the reference itself is single-threaded). */
EXPORT void orc_csr_mul(int64_t nrow, const int64_t *rowptr, const int32_t *colidx,
                        const double *nzval, const double *v, double alpha, double beta,
                        double *y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nrow; ++i) {
        double acc = (beta == 0.0) ? 0.0 : beta * y[i];
        for (int64_t j = rowptr[i]; j < rowptr[i + 1]; ++j)
            acc += nzval[j] * (v[colidx[j]] * alpha);
        y[i] = acc;
    }
}

/* ---- Julia stdlib mul!(C, A'::Adjoint{CSC}, v, a, b) as called from scaling.jl:255:
 *      per column: tmp = sum_j A[j,col]*v[rv[j]] (storage order); C[col] = a*tmp + b*C[col] ---- */
EXPORT void orc_csc_mul_t(int64_t nrow, int64_t ncol, const int64_t *colptr,
                          const int64_t *rowval, const double *nzval, const double *v,
                          double alpha, double beta, double *y) {
    (void)nrow;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t c = 0; c < ncol; ++c) {
        double tmp = 0.0;
        for (int64_t j = colptr[c]; j < colptr[c + 1]; ++j) tmp += nzval[j] * v[rowval[j]];
        y[c] = (beta == 0.0) ? alpha * tmp : alpha * tmp + beta * y[c];
    }
}

/* Stable CSC -> CSR (row-major copy with ascending column order inside each row). */
EXPORT void orc_csc_to_csr(int64_t nrow, int64_t ncol, const int64_t *colptr,
                           const int64_t *rowval, const double *nzval, int64_t *rowptr,
                           int32_t *colidx, double *outval) {
    memset(rowptr, 0, (size_t)(nrow + 1) * sizeof(int64_t));
    int64_t nnz = colptr[ncol];
    for (int64_t j = 0; j < nnz; ++j) rowptr[rowval[j] + 1] += 1;
    for (int64_t i = 0; i < nrow; ++i) rowptr[i + 1] += rowptr[i];
    int64_t *cursor = (int64_t *)__builtin_malloc((size_t)nrow * sizeof(int64_t));
    memcpy(cursor, rowptr, (size_t)nrow * sizeof(int64_t));
    for (int64_t c = 0; c < ncol; ++c)
        for (int64_t j = colptr[c]; j < colptr[c + 1]; ++j) {
            int64_t p = cursor[rowval[j]]++;
            colidx[p] = (int32_t)c;
            outval[p] = nzval[j];
        }
    __builtin_free(cursor);
}

/* ---- mul.jl:50-77  mul!(y, A::CSC, x::SparseVector, alpha, beta): the reference loop, literally.
 *      beta != 1: fill!(y, 0) or rmul!(y, beta); alpha == 0: return; for jp over the stored x: for kp in column rvx[jp]:
 *      y[i] += alpha * nzvA[kp] * nzx   (Julia's n-ary * is (alpha*a)*x) ---- */
EXPORT void orc_spmspv(int64_t m, int64_t n, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                       const int64_t *xind, const double *xval, int64_t nx, double alpha, double beta, double *y) {
    (void)n;
    if (beta != 1.0) {
        if (beta == 0.0) memset(y, 0, (size_t)m * sizeof(double));
        else for (int64_t i = 0; i < m; ++i) y[i] *= beta;
    }
    if (alpha == 0.0) return;
    for (int64_t jp = 0; jp < nx; ++jp) {
        const double nzx = xval[jp];
        const int64_t j = xind[jp];
        for (int64_t kp = colptr[j]; kp < colptr[j + 1]; ++kp) {
            const int64_t i = rowval[kp];
            y[i] += (alpha * nzval[kp]) * nzx;
        }
    }
}

/* ---- mul.jl:82-114  mul!(C::Dense, A::CSC, B::CSC, alpha, beta): the reference triple loop, literally (C column-major) ---- */
EXPORT void orc_spgemm_dense(int64_t m, int64_t p, const int64_t *acolptr, const int64_t *arowval, const double *anzval,
                             const int64_t *bcolptr, const int64_t *browval, const double *bnzval, double alpha, double beta,
                             double *C) {
    if (beta != 1.0) {
        if (beta != 0.0) for (int64_t i = 0; i < m * p; ++i) C[i] *= beta;
        else memset(C, 0, (size_t)(m * p) * sizeof(double));
    }
    for (int64_t col = 0; col < p; ++col)
        for (int64_t jp = bcolptr[col]; jp < bcolptr[col + 1]; ++jp) {
            const double nzB = bnzval[jp];
            const int64_t j = browval[jp];
            for (int64_t kp = acolptr[j]; kp < acolptr[j + 1]; ++kp) {
                const int64_t row = arowval[kp];
                C[row + col * m] += (alpha * anzval[kp]) * nzB;
            }
        }
}

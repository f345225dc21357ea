/*
 * synth_twin.c — HOST TWIN of the benchmark's synthetic count generator (test infrastructure, NOT the product).
 *
 * bench.py's `--impl reference` arm must build the SAME input as the CUDA arm without loading the product library
 * ("input built without loading libsevero_b200.so"). The generator of severo.jl_b200/csrc/synth.cu is defined
 * from integer Philox4x32-10 plus individually rounded IEEE Float64 operations (no FMA, a fixed-order exp), and its gene /
 * cell tables come from host libm in both programs, so this file reproduces it bit for bit when compiled with
 * -ffp-contract=off (tests/test_gpu_synth_twin.py compares the two on the GPU box).
 *
 * Besides the columns themselves it offers the one streaming pass the reference arm needs over ALL genes — library sizes
 * (normalize.jl:24), per-gene moments of the raw counts in the reference's sequential Welford order (scaling.jl:18-34) and a
 * per-gene histogram of the count values, from which standardized_var_clipped (variablefeatures.jl:19-28) is a short sum —
 * so that the 2.5e9-nonzero full matrix never has to be held on the host.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

static inline void philox4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* the fixed-order exp of synth.cu (det_exp): same operations, same order, no contraction */
static inline double det_exp(double x) {
    const double kf = floor(x * 1.4426950408889634 + 0.5);
    const double r = (x - kf * 6.93147180369123816490e-01) - kf * 1.90821492927058770002e-10;
    double p = 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    const int k = (int)kf;
    if (k < -1021) return 0.0;
    union { uint64_t u; double d; } s;
    s.u = (uint64_t)(k + 1023) << 52;
    return p * s.d;
}

static inline int poisson_count(double lam, uint32_t bits) {
    if (lam > 600.0) lam = 600.0; /* keeps p0 = exp(-lam) a normal number; part of the generator's definition */
    const double u = ((double)bits + 0.5) * (1.0 / 4294967296.0);
    if (u < 1.0 - lam) return 0;
    double p = det_exp(-lam);
    if (u < p) return 0;
    double cdf = p;
    int k = 0;
    while (u >= cdf && k < 4096) {
        ++k;
        p = p * (lam / (double)k);
        cdf = cdf + p;
        if (p < 1e-300) break;
    }
    return k;
}

static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return ((double)(splitmix64(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline double gauss(uint64_t *s) {
    const double u1 = u01(s), u2 = u01(s);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

/* gene tables of svb_synth_counts: lamtab[j*K + c] = (scale * p_j) * (fold if program c raises gene j else 1). Returns the
 * calibrated scale (0 on allocation failure). */
EXPORT double orc_synth_gene_tables(int64_t genes, int K, double mean_nnz_per_cell, double fold, uint64_t seed, double *lamtab) {
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 12345;
    double *p = (double *)malloc((size_t)genes * sizeof(double));
    uint8_t *up = (uint8_t *)calloc((size_t)genes * (size_t)K, 1);
    if (!p || !up) { free(p); free(up); return 0.0; }
    double psum = 0.0;
    for (int64_t j = 0; j < genes; ++j) {
        p[j] = exp(1.8 * gauss(&s));
        psum += p[j];
    }
    for (int64_t j = 0; j < genes; ++j) p[j] /= psum;
    for (int64_t j = 0; j < genes; ++j)
        for (int c = 0; c < K; ++c)
            if (u01(&s) < 0.05) up[(size_t)j * K + c] = 1;
    double lo = 1.0, hi = 1e9;
    for (int it = 0; it < 200; ++it) {
        const double mid = sqrt(lo * hi);
        double tot = 0.0;
        for (int64_t j = 0; j < genes; ++j) {
            const double base = mid * p[j];
            int nup = 0;
            for (int c = 0; c < K; ++c) nup += up[(size_t)j * K + c];
            const double nz = (K - nup) * (1.0 - exp(-base)) + nup * (1.0 - exp(-base * fold));
            tot += nz / K;
        }
        if (tot < mean_nnz_per_cell) lo = mid; else hi = mid;
    }
    const double scale = sqrt(lo * hi);
    for (int64_t j = 0; j < genes; ++j) {
        const double base = scale * p[j];
        for (int c = 0; c < K; ++c) lamtab[(size_t)j * K + c] = up[(size_t)j * K + c] ? base * fold : base;
    }
    free(p);
    free(up);
    return scale;
}

/* cell parameters of svb_synth_counts for the cells [row0, row0 + rows) */
EXPORT void orc_synth_cell_params(int64_t row0, int64_t rows, int K, uint64_t seed, double *lib, uint8_t *prog) {
    const double sigma_l = 0.35;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        const uint64_t i = (uint64_t)(row0 + r);
        uint32_t u[4];
        philox4((uint32_t)i, (uint32_t)(i >> 32), 0xC0FFEEu, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), u);
        const double u1 = ((double)u[0] + 0.5) * (1.0 / 4294967296.0);
        const double u2 = ((double)u[1] + 0.5) * (1.0 / 4294967296.0);
        const double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        lib[r] = exp(sigma_l * z - 0.5 * sigma_l * sigma_l);
        prog[r] = (uint8_t)(u[2] % (uint32_t)K);
    }
}

/* The counts of gene g for a block of cells [r, r + nb) (r a multiple of 4, nb <= GB): three phases so that the first two
 * vectorise — Philox for every cell quad; lam and the pre-filter u < 1 - lam for every cell; the sampler proper only for the
 * few cells that pass. Writes x[0..nb) (0 for the filtered cells). Same arithmetic per pair as poisson_count alone. */
#define GB 1024
static inline void block_counts(int64_t row0, int64_t r, int nb, int64_t g, uint64_t seed, const double *lib, const uint8_t *prog,
                                const double *lg, uint32_t *bits, int32_t *x) {
    const int nq = (nb + 3) / 4;
    const uint64_t q0 = (uint64_t)(row0 + r) >> 2;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int qi = 0; qi < nq; ++qi) {
        const uint64_t q = q0 + (uint64_t)qi;
        uint32_t c0 = (uint32_t)q, c1 = (uint32_t)(q >> 32), c2 = (uint32_t)g, c3 = 2u, a0 = k0, a1 = k1;
        for (int rd = 0; rd < 10; ++rd) {
            const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
            const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a0;
            const uint32_t n1 = (uint32_t)p1;
            const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ a1;
            const uint32_t n3 = (uint32_t)p0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a0 += 0x9E3779B9u;
            a1 += 0xBB67AE85u;
        }
        bits[4 * qi + 0] = c0; bits[4 * qi + 1] = c1; bits[4 * qi + 2] = c2; bits[4 * qi + 3] = c3;
    }
    for (int i = 0; i < nb; ++i) {
        const double lam = lib[r + i] * lg[prog[r + i]];
        const double u = ((double)bits[i] + 0.5) * (1.0 / 4294967296.0);
        x[i] = !(u < 1.0 - lam); /* candidates (NaN-safe: the same sense as the sampler's test) */
    }
    for (int i = 0; i < nb; ++i)
        if (x[i]) x[i] = poisson_count(lib[r + i] * lg[prog[r + i]], bits[i]);
}

/* One streaming pass over ALL genes (row0 must be a multiple of 4, as in the device generator):
 *   libsize[rows]   normalize.jl:24 row sums (exact integers)
 *   gene_nnz[genes] stored entries per gene
 *   mean/var[genes] scaling.jl:18-34 sequential Welford over the stored counts, count pre-seeded with the implicit zeros —
 *                   needs the gene's nnz first, hence two sweeps of the gene (the second from a per-thread buffer)
 *   hist[genes*HB]  hist[g*HB + c] = #cells with count c (1 <= c < HB); *overflow = #entries with count >= HB
 * Genes are independent: OpenMP over genes changes no bit. */
EXPORT int orc_synth_pass_stats(int64_t row0, int64_t rows, int64_t genes, int K, uint64_t seed, const double *lamtab,
                                const double *lib, const uint8_t *prog, int HB, int64_t *libsize, int64_t *gene_nnz,
                                double *mean, double *var, int64_t *hist, int64_t *overflow) {
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    int64_t *libs = (int64_t *)calloc((size_t)nthreads * (size_t)rows, sizeof(int64_t));
    if (!libs) return -1;
    int64_t over = 0;
    int fail = 0;
#pragma omp parallel reduction(+ : over)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        int64_t *mylib = libs + (size_t)tid * (size_t)rows;
        int32_t *buf = (int32_t *)malloc((size_t)rows * sizeof(int32_t)); /* the gene's stored counts, in cell order */
        if (!buf) {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp for schedule(dynamic, 4)
        for (int64_t g = 0; g < genes; ++g) {
            if (!buf) continue;
            const double *lg = lamtab + (size_t)g * K;
            int64_t *hg = hist + (size_t)g * HB;
            int64_t nnz = 0;
            for (int64_t r = 0; r < rows; r += GB) {
                const int nb = (rows - r) < GB ? (int)(rows - r) : GB;
                uint32_t bits[GB];
                int32_t x[GB];
                block_counts(row0, r, nb, g, seed, lib, prog, lg, bits, x);
                for (int e = 0; e < nb; ++e) {
                    if (x[e]) {
                        buf[nnz++] = x[e];
                        mylib[r + e] += x[e];
                        if (x[e] < HB) hg[x[e]] += 1; else over += 1;
                    }
                }
            }
            gene_nnz[g] = nnz;
            int64_t count = rows - nnz; /* scaling.jl:21 */
            double mu = 0.0, s = 0.0;
            for (int64_t k = 0; k < nnz; ++k) {
                count += 1;
                const double xv = (double)buf[k];
                const double delta = xv - mu;
                mu += delta / (double)count;
                s += delta * (xv - mu);
            }
            mean[g] = mu;
            var[g] = s / (double)(rows - 1);
        }
        free(buf);
    }
    memset(libsize, 0, (size_t)rows * sizeof(int64_t));
    for (int t = 0; t < nthreads; ++t)
        for (int64_t r = 0; r < rows; ++r) libsize[r] += libs[(size_t)t * (size_t)rows + r];
    free(libs);
    *overflow = over;
    return fail ? -1 : 0;
}

/* the columns sel[0..nsel) as CSC (cells ascending inside a gene): pass 0 counts, pass 1 fills */
EXPORT void orc_synth_columns_count(int64_t row0, int64_t rows, int K, uint64_t seed, const double *lamtab, const double *lib,
                                    const uint8_t *prog, const int64_t *sel, int64_t nsel, int64_t *colnnz) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t c = 0; c < nsel; ++c) {
        const int64_t g = sel[c];
        const double *lg = lamtab + (size_t)g * K;
        int64_t nnz = 0;
        for (int64_t r = 0; r < rows; r += GB) {
            const int nb = (rows - r) < GB ? (int)(rows - r) : GB;
            uint32_t bits[GB];
            int32_t x[GB];
            block_counts(row0, r, nb, g, seed, lib, prog, lg, bits, x);
            for (int e = 0; e < nb; ++e) nnz += (x[e] != 0);
        }
        colnnz[c] = nnz;
    }
}

EXPORT void orc_synth_columns_fill(int64_t row0, int64_t rows, int K, uint64_t seed, const double *lamtab, const double *lib,
                                   const uint8_t *prog, const int64_t *sel, int64_t nsel, const int64_t *colptr,
                                   int64_t *rowval, int32_t *val) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t c = 0; c < nsel; ++c) {
        const int64_t g = sel[c];
        const double *lg = lamtab + (size_t)g * K;
        int64_t pos = colptr[c];
        for (int64_t r = 0; r < rows; r += GB) {
            const int nb = (rows - r) < GB ? (int)(rows - r) : GB;
            uint32_t bits[GB];
            int32_t x[GB];
            block_counts(row0, r, nb, g, seed, lib, prog, lg, bits, x);
            for (int e = 0; e < nb; ++e)
                if (x[e]) {
                    rowval[pos] = r + e;
                    val[pos] = x[e];
                    ++pos;
                }
        }
    }
}

/* variablefeatures.jl:19-28 from the histogram of a gene's counts: sum over the distinct values c of
 * hist[c] * min((c - mu)/sd, vmax)^2, plus the implicit zeros, / (m - 1); long-double accumulation like
 * orc_stdvar_clipped_i64 (the referee every summation order must be within 1 ulp of). */
EXPORT void orc_stdvar_clipped_hist(int64_t nrow, int64_t genes, int HB, const int64_t *hist, const int64_t *gene_nnz,
                                    const double *mu, const double *sd, double vmax, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < genes; ++g) {
        out[g] = 0.0;
        if (sd[g] == 0.0) continue;
        long double acc = 0.0L;
        for (int c = 1; c < HB; ++c) {
            const int64_t h = hist[(size_t)g * HB + c];
            if (!h) continue;
            double z = ((double)c - mu[g]) / sd[g];
            if (z > vmax) z = vmax;
            acc += (long double)h * (long double)(z * z);
        }
        double z0 = (0.0 - mu[g]) / sd[g];
        if (z0 > vmax) z0 = vmax;
        const double zterm = (double)(nrow - gene_nnz[g]) * (z0 * z0);
        out[g] = ((double)acc + zterm) / (double)(nrow - 1);
    }
}

/* normalize.jl:25-29,36 on int32 counts with caller-supplied library sizes (the HVG columns of a larger matrix) */
EXPORT void orc_lognorm_i32(int64_t nnz, const int64_t *rowval, const int32_t *nzval, const int64_t *s, double scale_factor,
                            double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < nnz; ++j) {
        const double t = scale_factor * (double)nzval[j];
        out[j] = log1p(t / (double)s[rowval[j]]);
    }
}

// synth.cu — synthetic single-cell count matrices for the benchmark configurations:
// Poisson counts x_ij ~ Poisson(L_i * lam_{j,c(i)}), lam_{j,c} = scale * p_j * f_{c,j}, with log-normal library factors L_i,
// heavy-tailed gene propensities p_j and K planted cell programs (fold-change on ~5% of the genes each) so that
// sigma_nu / sigma_{nu+1} has a gap (SURVEY trap T1). Counter-based Philox keyed on the GLOBAL (cell, gene) pair: any rank
// regenerates exactly its own rows, independent of the sharding.
//
// The generator is DEFINED so that a host program reproduces it bit for bit (round 2: the CPU reference arm of bench.py
// builds the same input on the host cores without loading this library, from a plain-C twin kept with the test infrastructure):
//   * gene tables (p_j, f, the calibrated scale) and cell parameters (L_i, c(i)) are computed on the HOST in Float64 with
//     libm (identical for both programs of one box) and uploaded;
//   * everything per (cell, gene) pair uses integer Philox plus individually rounded IEEE Float64 operations only
//     (__dmul_rn / __dadd_rn / __ddiv_rn: no FMA contraction), including a small fixed-order exp (det_exp below):
//       lam = L_i * lamtab[j][c(i)] ; u = (bits + 0.5) * 2^-32 ;
//       x = 0 if u < 1 - lam (cheap exact pre-filter: 1 - lam <= exp(-lam)), else the inverse-CDF walk from p0 = det_exp(-lam).
// Output: CSC cells x genes, int32 counts, rows ascending inside every gene.
#include "svb_internal.h"
#include "layout.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

using namespace svb;

namespace svb {

__host__ __device__ __forceinline__ void philox4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                                 uint32_t out[4]) {
#pragma unroll
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

// exp(x) for x <= 0 from individually rounded multiplies / adds in a fixed order (range reduction by ln 2, Taylor polynomial
// of degree 13 in Horner form, exact scaling by 2^k): ~1e-16 relative accuracy, and — the point — the same bits on the
// device and in the host twin, which CUDA's and glibc's exp() do not give.
__device__ __forceinline__ double det_exp(double x) {
    const double kf = floor(__dadd_rn(__dmul_rn(x, 1.4426950408889634), 0.5));
    const double r = __dsub_rn(__dsub_rn(x, __dmul_rn(kf, 6.93147180369123816490e-01)), __dmul_rn(kf, 1.90821492927058770002e-10));
    double p = 1.0 / 6227020800.0;
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 479001600.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 39916800.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 3628800.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 362880.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 40320.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 5040.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 720.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 120.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 24.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 6.0);
    p = __dadd_rn(__dmul_rn(p, r), 0.5);
    p = __dadd_rn(__dmul_rn(p, r), 1.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0);
    const int k = (int)kf;
    if (k < -1021) return 0.0;
    return __dmul_rn(p, __longlong_as_double((long long)(k + 1023) << 52));
}

__device__ __forceinline__ int poisson_count(double lam, uint32_t bits) {
    if (lam > 600.0) lam = 600.0;  // keeps p0 = exp(-lam) a normal number; part of the generator's definition
    const double u = __dmul_rn(__dadd_rn((double)bits, 0.5), 1.0 / 4294967296.0);
    if (u < __dsub_rn(1.0, lam)) return 0;  // P(X = 0) = exp(-lam) >= 1 - lam
    double p = det_exp(-lam);
    if (u < p) return 0;
    double cdf = p;
    int k = 0;
    while (u >= cdf && k < 4096) {
        ++k;
        p = __dmul_rn(p, __ddiv_rn(lam, (double)k));
        cdf = __dadd_rn(cdf, p);
        if (p < 1e-300) break;
    }
    return k;
}

// grid (nchunks, genes); block 256 threads, thread t owns cells 4t..4t+3 of the 1024-cell chunk.
// PASS 0: cnt[g * nchunks + chunk] = #nonzeros ; PASS 1: write them at off[g * nchunks + chunk] + rank.
template <int PASS>
__global__ void __launch_bounds__(256) synth_kernel(int64_t row0, int64_t rows, int64_t genes, uint64_t seed,
                                                    const double *__restrict__ lib, const uint8_t *__restrict__ prog,
                                                    const double *__restrict__ lamtab, int programs,
                                                    int64_t nchunks, int64_t *__restrict__ cnt, int32_t *__restrict__ rowidx,
                                                    int32_t *__restrict__ val) {
    __shared__ int warp_cnt[8];
    const int64_t chunk = blockIdx.x, g = blockIdx.y;
    const int64_t r0 = chunk * 1024 + (int64_t)threadIdx.x * 4;
    const double *lg = lamtab + g * programs;
    int x[4] = {0, 0, 0, 0};
    if (r0 < rows) {
        const uint64_t q = (uint64_t)(row0 + r0) >> 2;  // cell quad (row0 is a multiple of 4 by construction)
        uint32_t u[4];
        philox4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)g, 2u, (uint32_t)seed, (uint32_t)(seed >> 32), u);
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (r0 + e < rows) x[e] = poisson_count(__dmul_rn(lib[r0 + e], __ldg(lg + prog[r0 + e])), u[e]);
    }
    const int mine = (x[0] != 0) + (x[1] != 0) + (x[2] != 0) + (x[3] != 0);
    // block exclusive scan of `mine`
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_cnt[wid] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < wid) wbase += warp_cnt[w];
        total += warp_cnt[w];
    }
    if (PASS == 0) {
        if (threadIdx.x == 0) cnt[g * nchunks + chunk] = total;
    } else {
        int64_t pos = cnt[g * nchunks + chunk] + wbase + incl - mine;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (x[e] != 0) {
                rowidx[pos] = (int32_t)(r0 + e);
                val[pos] = x[e];
                ++pos;
            }
        }
    }
}

static inline uint64_t splitmix64(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t &s) { return ((double)(splitmix64(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline double gauss(uint64_t &s) {
    const double u1 = u01(s), u2 = u01(s);
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
}

}  // namespace svb

extern "C" int svb_synth_counts(int64_t m_total, int64_t genes, int64_t row0, int64_t row1, double mean_nnz_per_cell,
                                int64_t programs, double fold, uint64_t seed, svb_matrix_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(out, SVB_EARG, "svb_synth_counts: null argument");
    SVB_CHECK(m_total >= 1 && genes >= 1 && 0 <= row0 && row0 <= row1 && row1 <= m_total, SVB_EDIM, "svb_synth_counts: bad shape");
    SVB_CHECK(row0 % 4 == 0, SVB_EDIM, "svb_synth_counts: shard start must be a multiple of 4");
    SVB_CHECK(programs >= 1 && programs <= 255, SVB_EARG, "svb_synth_counts: 1 <= programs <= 255");
    SVB_CHECK(genes <= 65535, SVB_EDIM, "svb_synth_counts: genes must be <= 65535 (grid.y)");
    SVB_CHECK(mean_nnz_per_cell > 0 && mean_nnz_per_cell < 0.9 * genes, SVB_EARG, "svb_synth_counts: bad density");
    cudaStream_t st = ctx().stream;
    const int64_t rows = row1 - row0;
    const int K = (int)programs;
    // gene-level parameters on the host (same on every rank: keyed by the seed only)
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 12345;
    std::vector<double> p((size_t)genes);
    double psum = 0.0;
    for (int64_t j = 0; j < genes; ++j) {
        p[j] = std::exp(1.8 * gauss(s));
        psum += p[j];
    }
    for (auto &v : p) v /= psum;
    std::vector<uint8_t> up((size_t)genes * K, 0);  // program c raises gene j by `fold`
    for (int64_t j = 0; j < genes; ++j)
        for (int c = 0; c < K; ++c)
            if (u01(s) < 0.05) up[(size_t)j * K + c] = 1;
    // calibrate the global intensity so that E[#nonzeros per cell] = mean_nnz_per_cell (L = 1, averaged over programs)
    const double sigma_l = 0.35;
    auto expected_nnz = [&](double scale) {
        double tot = 0.0;
        for (int64_t j = 0; j < genes; ++j) {
            const double base = scale * p[j];
            int nup = 0;
            for (int c = 0; c < K; ++c) nup += up[(size_t)j * K + c];
            const double nz = (K - nup) * (1.0 - std::exp(-base)) + nup * (1.0 - std::exp(-base * fold));
            tot += nz / K;
        }
        return tot;
    };
    double lo = 1.0, hi = 1e9;
    for (int it = 0; it < 200; ++it) {
        const double mid = std::sqrt(lo * hi);
        if (expected_nnz(mid) < mean_nnz_per_cell) lo = mid; else hi = mid;
    }
    const double scale = std::sqrt(lo * hi);
    std::vector<double> lamtab((size_t)genes * K);  // lam_{j,c} = (scale * p_j) * f_{c,j}
    double lam_max = 0.0;
    for (int64_t j = 0; j < genes; ++j) {
        const double base = scale * p[j];
        for (int c = 0; c < K; ++c) {
            const double v = up[(size_t)j * K + c] ? base * fold : base;
            lamtab[(size_t)j * K + c] = v;
            lam_max = std::max(lam_max, v);
        }
    }
    (void)lam_max;  // intensities above 600 per (cell, gene) pair are clamped inside the sampler
    // cell parameters on the host, Float64 + libm: L_i = exp(sigma z - sigma^2/2), z from Box-Muller on two 32-bit uniforms
    std::vector<double> lib((size_t)std::max<int64_t>(rows, 1));
    std::vector<uint8_t> prog((size_t)std::max<int64_t>(rows, 1));
    for (int64_t r = 0; r < rows; ++r) {
        const uint64_t i = (uint64_t)(row0 + r);
        uint32_t u[4];
        philox4((uint32_t)i, (uint32_t)(i >> 32), 0xC0FFEEu, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), u);
        const double u1 = ((double)u[0] + 0.5) * (1.0 / 4294967296.0);
        const double u2 = ((double)u[1] + 0.5) * (1.0 / 4294967296.0);
        const double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
        lib[(size_t)r] = std::exp(sigma_l * z - 0.5 * sigma_l * sigma_l);
        prog[(size_t)r] = (uint8_t)(u[2] % (uint32_t)K);
    }

    DevBuf<double> d_lib((size_t)std::max<int64_t>(rows, 1)), d_lam((size_t)genes * K);
    DevBuf<uint8_t> d_prog((size_t)std::max<int64_t>(rows, 1));
    SVB_CUDA(cudaMemcpyAsync(d_lam.p, lamtab.data(), lamtab.size() * 8, cudaMemcpyHostToDevice, st));
    SVB_CUDA(cudaMemcpyAsync(d_lib.p, lib.data(), lib.size() * 8, cudaMemcpyHostToDevice, st));
    SVB_CUDA(cudaMemcpyAsync(d_prog.p, prog.data(), prog.size(), cudaMemcpyHostToDevice, st));
    const int64_t nchunks = std::max<int64_t>(1, (rows + 1023) / 1024);
    DevBuf<int64_t> cnt((size_t)(genes * nchunks + 1));
    SVB_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)(genes * nchunks + 1) * 8, st));
    svb_matrix_s *a = nullptr;
    if (rows > 0) {
        dim3 grid((unsigned)nchunks, (unsigned)genes);
        synth_kernel<0><<<grid, 256, 0, st>>>(row0, rows, genes, seed, d_lib.p, d_prog.p, d_lam.p, K, nchunks, cnt.p, nullptr, nullptr);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    exclusive_scan_i64(cnt.p, genes * nchunks + 1, st);
    int64_t nnz = 0;
    SVB_CUDA(cudaMemcpyAsync(&nnz, cnt.p + genes * nchunks, 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    a = matrix_alloc(rows, genes, nnz, SVB_I32);
    try {
        launch_strided_copy(cnt.p, nchunks, genes, a->colptr, st);
        SVB_CUDA(cudaMemcpyAsync(a->colptr + genes, cnt.p + genes * nchunks, 8, cudaMemcpyDeviceToDevice, st));
        if (rows > 0 && nnz > 0) {
            dim3 grid((unsigned)nchunks, (unsigned)genes);
            synth_kernel<1><<<grid, 256, 0, st>>>(row0, rows, genes, seed, d_lib.p, d_prog.p, d_lam.p, K, nchunks, cnt.p, a->rowidx,
                                                  (int32_t *)a->val);
            count_launch();
            SVB_LAUNCH_CHECK();
        }
        SVB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        delete a;
        throw;
    }
    *out = a;
    SVB_API_END
}

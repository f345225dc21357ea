// preprocess.cu — the pre-processing sweeps of the path, on the reference's own CSC layout:
//   svb_row_sums / svb_normalize : normalize.jl:17-38  (library size, sf*x/s, log1p)
//   svb_mean_var                 : scaling.jl:18-34,132-142 (sequential Welford per gene, order-exact)
//   svb_stdvar_clipped           : variablefeatures.jl:19-28
//   svb_scale                    : scaling.jl:199-217 (mean/std, x/std, upper clip at scale_max + mean/std)
// Arithmetic that must be bit-identical to Julia uses the _rn intrinsics so nvcc cannot contract
// a*b+c into an FMA (Julia does not).
#include "svb_internal.h"
#include <chrono>

#include <algorithm>
#include <cmath>
#include <numeric>

using namespace svb;

namespace svb {

static inline unsigned grid1(int64_t n, int threads = 256) {
    int64_t b = (n + threads - 1) / threads;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(b, 148 * 16));
}

// ---- library sizes ----------------------------------------------------------------------------
__global__ void row_sums_kernel(const int32_t *__restrict__ rowidx, const int32_t *__restrict__ val, int64_t nnz,
                                unsigned long long *__restrict__ s) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&s[rowidx[i]], (unsigned long long)(long long)val[i]);  // exact integer sum, order-free
}

// normalize.jl:27  nzB[j] = scale_factor * nzA[j] / s[rv[j]]  (left to right), :36 log1p
template <typename T>
__global__ void libnorm_kernel(const int32_t *__restrict__ rowidx, const int32_t *__restrict__ val, int64_t nnz,
                                 const long long *__restrict__ s, T sf, int do_log, T *__restrict__ out);

template <>
__global__ void libnorm_kernel<double>(const int32_t *__restrict__ rowidx, const int32_t *__restrict__ val, int64_t nnz,
                                         const long long *__restrict__ s, double sf, int do_log, double *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const double t = __dmul_rn(sf, (double)val[i]);
        const double v = __ddiv_rn(t, (double)s[rowidx[i]]);
        out[i] = do_log ? log1p(v) : v;
    }
}

template <>
__global__ void libnorm_kernel<float>(const int32_t *__restrict__ rowidx, const int32_t *__restrict__ val, int64_t nnz,
                                        const long long *__restrict__ s, float sf, int do_log, float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const float t = __fmul_rn(sf, (float)val[i]);
        const float v = __fdiv_rn(t, (float)s[rowidx[i]]);
        out[i] = do_log ? log1pf(v) : v;
    }
}

// ---- order-exact Welford: one thread walks one gene; genes are handed out longest-first so the
// lanes of a warp carry chains of similar length. Loads run four elements ahead of the chain. -------
template <typename T> __device__ __forceinline__ T sub_rn(T a, T b);
template <> __device__ __forceinline__ double sub_rn<double>(double a, double b) { return __dsub_rn(a, b); }
template <> __device__ __forceinline__ float sub_rn<float>(float a, float b) { return __fsub_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T div_rn(T a, T b);
template <> __device__ __forceinline__ double div_rn<double>(double a, double b) { return __ddiv_rn(a, b); }
template <> __device__ __forceinline__ float div_rn<float>(float a, float b) { return __fdiv_rn(a, b); }

// One Welford step (scaling.jl:26-29) in Float64 with the divide taken off the dependent chain: y = RN(1/count) is
// computed ahead (count is just a counter), then q0 = RN(delta*y), r = delta - count*q0 (exact, FMA), q = RN(q0 + r*y).
// With a correctly rounded reciprocal and a divisor whose significand is not all ones (an integer < 2^53 never is)
// this is the correctly rounded quotient (Markstein); checked against `/` on 6.4e8 random + adversarial pairs
// (see DESIGN.md). The dependent chain per element is 5 fp64 ops instead of a ~40-instruction software divide.
// Zero / huge / tiny deltas take the plain IEEE divide so signed zeros, infinities and subnormals stay exact.
// (kept out of line: inlined, the compiler evaluates the ~15-instruction divide for EVERY element and selects afterwards —
// ncu on the dense-gene case: 63 instructions per element on a single warp that issues one every ~3 cycles)
__device__ __noinline__ double welford_slow_div(double delta, double c) { return __ddiv_rn(delta, c); }

__device__ __forceinline__ void welford_step_f64(double v, double c, double y, double &mu, double &s) {
    const double delta = __dsub_rn(v, mu);
    const double q0 = __dmul_rn(delta, y);
    const double r = __fma_rn(-c, q0, delta);
    double q = __fma_rn(r, y, q0);
    // fast path for 2^-962 <= |delta| < 2^963 (biased exponent 61..1985), tested on the exponent bits with integer instructions
    const unsigned be = ((unsigned)__double2hiint(delta) >> 20) & 0x7ffu;
    if (__builtin_expect(be - 61u > 1924u, 0)) q = welford_slow_div(delta, c);
    mu = __dadd_rn(mu, q);
    s = __dadd_rn(s, __dmul_rn(delta, __dsub_rn(v, mu)));
}

// the same step with the fast quotient only; `bad` collects the range test (the caller redoes the steps exactly if it is set)
__device__ __forceinline__ void welford_step_fast(double v, double c, double y, double &mu, double &s, unsigned &bad) {
    const double delta = __dsub_rn(v, mu);
    const double q0 = __dmul_rn(delta, y);
    const double r = __fma_rn(-c, q0, delta);
    const double q = __fma_rn(r, y, q0);
    const unsigned be = ((unsigned)__double2hiint(delta) >> 20) & 0x7ffu;
    bad |= (unsigned)(be - 61u > 1924u);
    mu = __dadd_rn(mu, q);
    s = __dadd_rn(s, __dmul_rn(delta, __dsub_rn(v, mu)));
}

// RN(1/count) for the running counter. The lane that walks a dense gene is bound by instruction issue (one warp,
// fp64 at half rate): a correctly rounded reciprocal from scratch is ~35 fp64 instructions per element. Since the
// counter only increments, 1/(c+1) follows from 1/c by three FMA-Newton steps (relative error 1/c -> 1/c^2 ->
// 1/c^4 < 1 ulp for c >= 2^14; the third step is the Markstein correction that rounds correctly) — verified equal
// to RN(1/c) for every c in [16385, 6e7] on the CPU. Below 2^14 the exact reciprocal is used.
template <typename T> struct WelfordStep {
    __device__ __forceinline__ bool can4(long long) const { return false; }
    __device__ __forceinline__ void run4(T, T, T, T, long long, T &, T &) {}
    __device__ __forceinline__ void run(T v, long long count, T &mu, T &s) {
        const T delta = sub_rn<T>(v, mu);
        mu = add_rn<T>(mu, div_rn<T>(delta, (T)count));
        s = add_rn<T>(s, mul_rn<T>(delta, sub_rn<T>(v, mu)));
    }
};
template <> struct WelfordStep<double> {
    double y = 0.0;
    bool valid = false;
    __device__ __forceinline__ void run(double v, long long count, double &mu, double &s) {
        const double c = (double)count;
        if (valid && count > 16384) {
            double e = __fma_rn(-c, y, 1.0);
            y = __fma_rn(y, e, y);
            e = __fma_rn(-c, y, 1.0);
            y = __fma_rn(y, e, y);
            e = __fma_rn(-c, y, 1.0);
            y = __fma_rn(y, e, y);
        } else {
            y = __drcp_rn(c);
            valid = true;
        }
        welford_step_f64(v, c, y, mu, s);
    }
    // FOUR steps at once: the reciprocals of count+1 .. count+4 each by three Newton steps started at y = RN(1/count) —
    // four independent chains next to the data chain instead of one 6-FMA chain in front of every step (the dependent path
    // per element goes from 11 fp64 operations to 5: sub, mul, 2 fma, add). Equal to RN(1/d) for every count >= 16384 and
    // d < 2^32, jumps up to 8: tools/studies/recip_check.c (exhaustive).
    __device__ __forceinline__ bool can4(long long count) const { return valid && count >= 16384; }
    __device__ __forceinline__ void run4(double v0, double v1, double v2, double v3, long long count, double &mu, double &s) {
        const double c = (double)count;
        const double c1 = c + 1.0, c2 = c + 2.0, c3 = c + 3.0, c4 = c + 4.0;
        double y1 = y, y2 = y, y3 = y, y4 = y;
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const double e1 = __fma_rn(-c1, y1, 1.0), e2 = __fma_rn(-c2, y2, 1.0), e3 = __fma_rn(-c3, y3, 1.0), e4 = __fma_rn(-c4, y4, 1.0);
            y1 = __fma_rn(y1, e1, y1);
            y2 = __fma_rn(y2, e2, y2);
            y3 = __fma_rn(y3, e3, y3);
            y4 = __fma_rn(y4, e4, y4);
        }
        // the four steps WITHOUT a branch between them (the s-updates of one step then overlap the mean chain of the next: a
        // single warp issues in order, and a branch per element kept the compiler from interleaving them — 80 ns per element);
        // the range test of the fast quotient is accumulated and, if any step failed it (rare), the four are redone exactly
        const double mu0 = mu, s0 = s;
        unsigned bad = 0;
        welford_step_fast(v0, c1, y1, mu, s, bad);
        welford_step_fast(v1, c2, y2, mu, s, bad);
        welford_step_fast(v2, c3, y3, mu, s, bad);
        welford_step_fast(v3, c4, y4, mu, s, bad);
        if (__builtin_expect(bad != 0u, 0)) {
            mu = mu0;
            s = s0;
            welford_step_f64(v0, c1, y1, mu, s);
            welford_step_f64(v1, c2, y2, mu, s);
            welford_step_f64(v2, c3, y3, mu, s);
            welford_step_f64(v3, c4, y4, mu, s);
        }
        y = y4;
    }
};

// Streaming part shared by both Welford kernels. One warp owns 32 genes (one per lane; genes are handed out
// longest-first, so the 32 chains have similar length). Per round the warp copies the next 32 values of each of
// its 32 genes global -> shared with cp.async (one fully coalesced 32-element row per gene, 32 copies in flight
// per lane, double-buffered so the next round streams in while the lanes run their dependent chains on the current
// one); rows are padded to 33 so both the copy (row g, column lane) and the chain reads (row lane, column i) are
// bank-conflict free. Measured at C3 (densest gene: 1.3 M stored values): 193 ms for the first version (each
// thread walking its own gene, software divide on the chain) -> 160 ms. What remains is the dependent fp64 chain
// itself, ~120 ns per element of the densest gene (sub, mul, 2 fma, add at fp64 latency, next to the 6-FMA
// reciprocal recurrence): an order-exact Welford cannot go below (length of the longest gene) x (chain latency).
template <typename VI>
__device__ __forceinline__ void cp_async_elem(VI *smem_dst, const VI *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (sizeof(VI) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}

constexpr int WF_WARPS = 2;  // warps per block

template <typename VI, typename T>
__device__ __forceinline__ void welford_warp_stream(const VI *__restrict__ val, int64_t beg, int64_t len, long long &count, T &mu,
                                                    T &s, VI (*buf)[32][33]) {
    const int lane = threadIdx.x & 31;
    WelfordStep<T> step;
    int64_t maxlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    auto issue = [&](int64_t base, int b) {
#pragma unroll 8
        for (int g = 0; g < 32; ++g) {
            const int64_t bg = __shfl_sync(0xffffffffu, beg, g), lg = __shfl_sync(0xffffffffu, len, g);
            const int64_t idx = base + lane;
            if (idx < lg) cp_async_elem<VI>(&buf[b][g][lane], val + bg + idx);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (maxlen > 0) issue(0, 0);
    int b = 0;
    for (int64_t base = 0; base < maxlen; base += 32, b ^= 1) {
        if (base + 32 < maxlen) {
            issue(base + 32, b ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        const int64_t rem = len - base;
        const int cnt = rem >= 32 ? 32 : (rem > 0 ? (int)rem : 0);
        int i = 0;
        for (; i + 4 <= cnt && step.can4(count); i += 4) {
            step.run4((T)buf[b][lane][i], (T)buf[b][lane][i + 1], (T)buf[b][lane][i + 2], (T)buf[b][lane][i + 3], count, mu, s);
            count += 4;
        }
        for (; i < cnt; ++i) {
            count += 1;
            step.run((T)buf[b][lane][i], count, mu, s);
        }
        __syncwarp();  // everybody is done with buffer b before the round after next overwrites it
    }
}

template <typename VI, typename T>
__global__ void __launch_bounds__(32 * WF_WARPS) welford_kernel(const int64_t *__restrict__ colptr, const VI *__restrict__ val,
                                                                const int32_t *__restrict__ order, int64_t ncol, int64_t nrow,
                                                                double *__restrict__ mu_out, double *__restrict__ var_out) {
    __shared__ VI buf[WF_WARPS][2][32][33];
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = g < ncol;
    const int64_t c = have ? order[g] : 0;
    const int64_t beg = have ? colptr[c] : 0, end = have ? colptr[c + 1] : 0;
    long long count = nrow - (end - beg);  // scaling.jl:21 implicit zeros first
    T mu = (T)0, s = (T)0;
    welford_warp_stream<VI, T>(val, beg, end - beg, count, mu, s, buf[threadIdx.x >> 5]);
    if (have) {
        mu_out[c] = (double)mu;
        var_out[c] = (double)div_rn<T>(s, (T)(nrow - 1));
    }
}

// Same chain, continued across cell shards: the state (count, mu, s) of every gene enters from the previous
// rank and leaves for the next one, so the bits equal those of ONE sequential pass over all cells.
template <typename VI>
__global__ void __launch_bounds__(32 * WF_WARPS) welford_carry_kernel(const int64_t *__restrict__ colptr, const VI *__restrict__ val,
                                                                      const int32_t *__restrict__ order, int64_t ncol,
                                                                      long long *__restrict__ count_io, double *__restrict__ mu_io,
                                                                      double *__restrict__ s_io) {
    __shared__ VI buf[WF_WARPS][2][32][33];
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = g < ncol;
    const int64_t c = have ? order[g] : 0;
    const int64_t beg = have ? colptr[c] : 0, end = have ? colptr[c + 1] : 0;
    long long count = have ? count_io[c] : 0;
    double mu = have ? mu_io[c] : 0.0, s = have ? s_io[c] : 0.0;
    welford_warp_stream<VI, double>(val, beg, end - beg, count, mu, s, buf[threadIdx.x >> 5]);
    if (have) {
        count_io[c] = count;
        mu_io[c] = mu;
        s_io[c] = s;
    }
}

// ---- standardized_var_clipped: one warp per gene, double-double accumulation -----------------------
__device__ __forceinline__ void two_sum(double a, double b, double &s, double &e) {
    s = __dadd_rn(a, b);
    const double bb = __dsub_rn(s, a);
    e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
}
__device__ __forceinline__ void dd_add(double &hi, double &lo, double x) {
    double s, e;
    two_sum(hi, x, s, e);
    lo = __dadd_rn(lo, e);
    hi = s;
}

__global__ void __launch_bounds__(256) stdvar_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ val,
                                                     const int32_t *__restrict__ order, int64_t ncol, int64_t nrow,
                                                     const double *__restrict__ mu, const double *__restrict__ sd, double vmax,
                                                     double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= ncol) return;
    const int64_t c = order[warp];
    const double m_c = mu[c], s_c = sd[c];
    if (s_c == 0.0) {  // variablefeatures.jl:24
        if (lane == 0) out[c] = 0.0;
        return;
    }
    const int64_t beg = colptr[c], end = colptr[c + 1];
    double hi = 0.0, lo = 0.0;
    for (int64_t k = beg + lane; k < end; k += 32) {
        double z = __ddiv_rn(__dsub_rn((double)val[k], m_c), s_c);
        z = fmin(z, vmax);
        dd_add(hi, lo, __dmul_rn(z, z));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ohi = __shfl_xor_sync(0xffffffffu, hi, o);
        const double olo = __shfl_xor_sync(0xffffffffu, lo, o);
        double s, e;
        two_sum(hi, ohi, s, e);
        lo = __dadd_rn(__dadd_rn(lo, olo), e);
        hi = s;
    }
    if (lane == 0) {
        const double acc = __dadd_rn(hi, lo);
        double z0 = __ddiv_rn(__dsub_rn(0.0, m_c), s_c);
        z0 = fmin(z0, vmax);
        const double zterm = __dmul_rn((double)(nrow - (end - beg)), __dmul_rn(z0, z0));
        out[c] = __ddiv_rn(__dadd_rn(acc, zterm), (double)(nrow - 1));
    }
}

// ---- scale_data --------------------------------------------------------------------------------------
// per gene: sd = sqrt(var); mu_s = mean/sd (stored mu, trap T3); smax = scale_max + mu_s
template <typename TO>
__global__ void scale_prepare_kernel(const double *__restrict__ mean, const double *__restrict__ var, int64_t ncol,
                                     double scale_max, double *__restrict__ sd, double *__restrict__ mu_s, double *__restrict__ smax) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const double s = sqrt(var[c]);
    sd[c] = s;
    // mu = zeros(R, d); mu[i] = mean (rounded to R); mu[i] /= std (scaling.jl:205-207)
    const TO m0 = (TO)mean[c];
    const TO m1 = (TO)__ddiv_rn((double)m0, s);
    mu_s[c] = (double)m1;
    smax[c] = (double)((TO)scale_max + m1);
}

template <typename VI, typename TO>
__global__ void scale_apply_kernel(const int64_t *__restrict__ colptr, const VI *__restrict__ val, const double *__restrict__ sd,
                                   const double *__restrict__ smax, TO *__restrict__ out) {
    const int64_t c = blockIdx.x;
    const int64_t beg = colptr[c], end = colptr[c + 1];
    const double s = sd[c], cap = smax[c];
    for (int64_t k = beg + (int64_t)blockIdx.y * blockDim.x + threadIdx.x; k < end; k += (int64_t)gridDim.y * blockDim.x) {
        const double v = __ddiv_rn((double)val[k], s);  // scaling.jl:211
        out[k] = (TO)((v > cap) ? cap : v);              // scaling.jl:212 (upper clip only)
    }
}

// float input: statistics and the division run in Float32 (scaling.jl:43-44, :211 with T = Float32)
__global__ void scale_apply_f32_kernel(const int64_t *__restrict__ colptr, const float *__restrict__ val,
                                       const double *__restrict__ sd, const double *__restrict__ smax, float *__restrict__ out) {
    const int64_t c = blockIdx.x;
    const int64_t beg = colptr[c], end = colptr[c + 1];
    const float s = (float)sd[c], cap = (float)smax[c];
    for (int64_t k = beg + (int64_t)blockIdx.y * blockDim.x + threadIdx.x; k < end; k += (int64_t)gridDim.y * blockDim.x) {
        const float v = __fdiv_rn(val[k], s);
        out[k] = (v > cap) ? cap : v;
    }
}

// columns ordered by decreasing length (host; ncol is the number of genes)
static void column_order(const svb_matrix_s *a, DevBuf<int32_t> &order) {
    cudaStream_t st = ctx().stream;
    std::vector<int64_t> cp((size_t)a->ncol + 1);
    SVB_CUDA(cudaMemcpyAsync(cp.data(), a->colptr, cp.size() * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    std::vector<int32_t> ord((size_t)a->ncol);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t x, int32_t y) { return cp[x + 1] - cp[x] > cp[y + 1] - cp[y]; });
    order.alloc(std::max<size_t>(ord.size(), 1));
    if (!ord.empty()) SVB_CUDA(cudaMemcpyAsync(order.p, ord.data(), ord.size() * 4, cudaMemcpyHostToDevice, st));
    SVB_CUDA(cudaStreamSynchronize(st));
}

static void mean_var_device(const svb_matrix_s *a, double *d_mu, double *d_var) {
    if (a->ncol == 0) return;
    DevBuf<int32_t> order;
    column_order(a, order);
    const unsigned grid = (unsigned)((a->ncol + 63) / 64);
    cudaStream_t st = ctx().stream;
    switch (a->vtype) {
        case SVB_I32: welford_kernel<int32_t, double><<<grid, 64, 0, st>>>(a->colptr, (const int32_t *)a->val, order.p, a->ncol, a->nrow, d_mu, d_var); break;
        case SVB_F64: welford_kernel<double, double><<<grid, 64, 0, st>>>(a->colptr, (const double *)a->val, order.p, a->ncol, a->nrow, d_mu, d_var); break;
        case SVB_F32: welford_kernel<float, float><<<grid, 64, 0, st>>>(a->colptr, (const float *)a->val, order.p, a->ncol, a->nrow, d_mu, d_var); break;
        default: throw Error(SVB_EARG, "bad vtype");
    }
    count_launch();
    SVB_LAUNCH_CHECK();
    SVB_CUDA(cudaStreamSynchronize(st));
}

// ---- upload + order-exact moments, pipelined (svb_csc_upload_lognorm_moments) ---------------------------------------------
// Y[k] = log1p((sf * c_k) / s[row_k]) for the entries of the listed columns only (the arithmetic of libnorm_kernel<double>)
__global__ void __launch_bounds__(256) lognorm_cols_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                           const int32_t *__restrict__ val, const int32_t *__restrict__ cols,
                                                           const long long *__restrict__ s, double sf, double *__restrict__ Y) {
    const int64_t c = cols[blockIdx.x];
    const int64_t b = colptr[c], e = colptr[c + 1];
    for (int64_t k = b + (int64_t)blockIdx.y * blockDim.x + threadIdx.x; k < e; k += (int64_t)gridDim.y * blockDim.x) {
        const double t = __dmul_rn(sf, (double)val[k]);
        Y[k] = log1p(__ddiv_rn(t, (double)s[rowidx[k]]));
    }
}

// rowidx of the listed columns from their staged host indices (Int64 or Int32, any base)
template <typename TH>
__global__ void __launch_bounds__(256) convert_cols_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ cols,
                                                           const int64_t *__restrict__ soff, const TH *__restrict__ stage, int64_t base,
                                                           int32_t *__restrict__ rowidx, int *__restrict__ overflow) {
    const int64_t c = cols[blockIdx.x];
    const int64_t b = colptr[c], len = colptr[c + 1] - b;
    const TH *src = stage + soff[blockIdx.x];
    for (int64_t k = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; k < len; k += (int64_t)gridDim.y * blockDim.x) {
        const int64_t v = (int64_t)src[k] - base;
        if (v > 2147483647LL || v < 0) *overflow = 1;
        rowidx[b + k] = (int32_t)v;
    }
}

template <typename TH>
static void upload_moments_pipeline(svb_matrix_s *a, const int64_t *h_colptr, const TH *h_rowval, const int32_t *h_counts, int base,
                                    const long long *d_lib, double sf, double *h_mean, double *h_var) {
    Context &C = ctx();
    cudaStream_t st = C.stream;
    const int64_t n = a->ncol, nnz = a->nnz;
    // columns by decreasing length: the longest Welford chains start first and run while the rest crosses PCIe
    std::vector<int32_t> ord((size_t)n);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t x, int32_t y) { return h_colptr[x + 1] - h_colptr[x] > h_colptr[y + 1] - h_colptr[y]; });
    constexpr int GW = 64;
    const int ng = (int)((n + GW - 1) / GW);
    std::vector<int64_t> soff((size_t)n);
    int64_t emax = 1;
    for (int g = 0; g < ng; ++g) {
        int64_t off = 0;
        for (int64_t i = (int64_t)g * GW; i < std::min<int64_t>(n, (int64_t)(g + 1) * GW); ++i) {
            soff[(size_t)i] = off;
            off += h_colptr[ord[(size_t)i] + 1] - h_colptr[ord[(size_t)i]];
        }
        emax = std::max(emax, off);
    }
    DevBuf<int32_t> d_cols((size_t)std::max<int64_t>(n, 1));
    DevBuf<int64_t> d_soff((size_t)std::max<int64_t>(n, 1));
    DevBuf<TH> stage0((size_t)emax), stage1((size_t)emax);
    DevBuf<double> Y((size_t)std::max<int64_t>(nnz, 1)), d_mean((size_t)std::max<int64_t>(n, 1)), d_var((size_t)std::max<int64_t>(n, 1));
    DevBuf<int> d_over(1);
    SVB_CUDA(cudaMemsetAsync(d_over.p, 0, sizeof(int), st));
    SVB_CUDA(cudaMemcpyAsync(d_cols.p, ord.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    SVB_CUDA(cudaMemcpyAsync(d_soff.p, soff.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    SVB_CUDA(cudaStreamSynchronize(st));  // colptr (uploaded by the caller on st), the lists and the flag are in place
    struct Streams {
        cudaStream_t copy = nullptr, conv = nullptr;
        std::vector<cudaStream_t> wf;
        std::vector<cudaEvent_t> ev, evc;
        ~Streams() {
            for (auto s_ : wf) if (s_) cudaStreamDestroy(s_);
            for (auto e_ : ev) if (e_) cudaEventDestroy(e_);
            for (auto e_ : evc) if (e_) cudaEventDestroy(e_);
            if (conv) cudaStreamDestroy(conv);
            if (copy) cudaStreamDestroy(copy);
        }
    } S;
    SVB_CUDA(cudaStreamCreateWithFlags(&S.copy, cudaStreamNonBlocking));
    SVB_CUDA(cudaStreamCreateWithFlags(&S.conv, cudaStreamNonBlocking));  // the row-index conversion: NOT on the copy stream (the copy engine idled behind it: 32 x 0.3 ms)
    const int nwf = std::min(ng, 16);
    S.wf.assign((size_t)nwf, nullptr);
    for (auto &s_ : S.wf) SVB_CUDA(cudaStreamCreateWithFlags(&s_, cudaStreamNonBlocking));
    S.ev.assign((size_t)ng, nullptr);
    for (auto &e_ : S.ev) SVB_CUDA(cudaEventCreateWithFlags(&e_, cudaEventDisableTiming));
    S.evc.assign((size_t)ng, nullptr);
    for (auto &e_ : S.evc) SVB_CUDA(cudaEventCreateWithFlags(&e_, cudaEventDisableTiming));
    const bool dbg = getenv("SVB_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    for (int g = 0; g < ng; ++g) {
        const int64_t i0 = (int64_t)g * GW, i1 = std::min<int64_t>(n, i0 + GW);
        const int ncg = (int)(i1 - i0);
        TH *stage = (g & 1) ? stage1.p : stage0.p;  // reused every other group: after the conversion of group g-2
        if (g >= 2) SVB_CUDA(cudaStreamWaitEvent(S.copy, S.evc[(size_t)(g - 2)], 0));
        for (int64_t i = i0; i < i1; ++i) {
            const int64_t c = ord[(size_t)i], b = h_colptr[c] - base, len = h_colptr[c + 1] - h_colptr[c];
            if (len == 0) continue;
            SVB_CUDA(cudaMemcpyAsync(stage + soff[(size_t)i], h_rowval + b, (size_t)len * sizeof(TH), cudaMemcpyHostToDevice, S.copy));
            SVB_CUDA(cudaMemcpyAsync((int32_t *)a->val + b, h_counts + b, (size_t)len * 4, cudaMemcpyHostToDevice, S.copy));
        }
        dim3 grid((unsigned)ncg, 8);
        SVB_CUDA(cudaEventRecord(S.ev[(size_t)g], S.copy));
        SVB_CUDA(cudaStreamWaitEvent(S.conv, S.ev[(size_t)g], 0));
        convert_cols_kernel<TH><<<grid, 256, 0, S.conv>>>(a->colptr, d_cols.p + i0, d_soff.p + i0, stage, base, a->rowidx, d_over.p);
        SVB_CUDA(cudaEventRecord(S.evc[(size_t)g], S.conv));
        cudaStream_t sw = S.wf[(size_t)(g % nwf)];
        SVB_CUDA(cudaStreamWaitEvent(sw, S.evc[(size_t)g], 0));
        lognorm_cols_kernel<<<grid, 256, 0, sw>>>(a->colptr, a->rowidx, (const int32_t *)a->val, d_cols.p + i0, d_lib, sf, Y.p);
        welford_kernel<double, double><<<(unsigned)((ncg + 63) / 64), 64, 0, sw>>>(a->colptr, Y.p, d_cols.p + i0, ncg, a->nrow, d_mean.p, d_var.p);
        count_launch(3);
        SVB_LAUNCH_CHECK();
    }
    const double t_issued = now();
    SVB_CUDA(cudaStreamSynchronize(S.copy));
    SVB_CUDA(cudaStreamSynchronize(S.conv));
    const double t_copied = now();
    for (auto s_ : S.wf) SVB_CUDA(cudaStreamSynchronize(s_));
    if (dbg)
        fprintf(stderr, "[svb upload+moments] %d groups: issued %.1f ms, copies + row conversion done %.1f ms, moments done %.1f ms\n", ng,
                (t_issued - t_begin) * 1e3, (t_copied - t_begin) * 1e3, (now() - t_begin) * 1e3);
    int over = 0;
    SVB_CUDA(cudaMemcpyAsync(&over, d_over.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaMemcpyAsync(h_mean, d_mean.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaMemcpyAsync(h_var, d_var.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_CHECK(!over, SVB_EDIM, "svb_csc_upload_lognorm_moments: a row index does not fit in int32 / lies below the base");
}

}  // namespace svb

extern "C" {

int svb_csc_upload_lognorm_moments(int64_t nrow, int64_t ncol, const int64_t *colptr, const void *rowval, int rowval_type,
                                   const int32_t *counts, int index_base, const int64_t *libsize, double scale_factor, double *mean,
                                   double *var, svb_matrix_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(out && colptr && libsize && mean && var, SVB_EARG, "svb_csc_upload_lognorm_moments: null argument");
    SVB_CHECK(nrow >= 1 && ncol >= 1 && nrow < 2147483647LL, SVB_EDIM, "svb_csc_upload_lognorm_moments: bad dimensions");
    SVB_CHECK(index_base == 0 || index_base == 1, SVB_EARG, "index_base must be 0 or 1");
    SVB_CHECK(rowval_type == SVB_I32 || rowval_type == SVB_I64, SVB_EARG, "rowval_type must be I32 or I64");
    const int64_t nnz = colptr[ncol] - index_base;
    SVB_CHECK(nnz >= 0 && colptr[0] == index_base, SVB_EDIM, "svb_csc_upload_lognorm_moments: malformed colptr");
    for (int64_t j = 0; j < ncol; ++j) SVB_CHECK(colptr[j + 1] >= colptr[j], SVB_EDIM, "svb_csc_upload_lognorm_moments: malformed colptr");
    SVB_CHECK(nnz == 0 || (rowval && counts), SVB_EARG, "svb_csc_upload_lognorm_moments: null rowval / counts");
    SVB_CHECK(scale_factor > 0.0, SVB_EARG, "svb_csc_upload_lognorm_moments: scale_factor must be positive");
    cudaStream_t st = ctx().stream;
    svb_matrix_s *a = matrix_alloc(nrow, ncol, nnz, SVB_I32);
    try {
        std::vector<int64_t> cp((size_t)ncol + 1);
        for (int64_t j = 0; j <= ncol; ++j) cp[(size_t)j] = colptr[j] - index_base;
        SVB_CUDA(cudaMemcpyAsync(a->colptr, cp.data(), cp.size() * 8, cudaMemcpyHostToDevice, st));
        DevBuf<long long> d_lib((size_t)nrow);
        SVB_CUDA(cudaMemcpyAsync(d_lib.p, libsize, (size_t)nrow * 8, cudaMemcpyHostToDevice, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        if (rowval_type == SVB_I64)
            upload_moments_pipeline<int64_t>(a, colptr, (const int64_t *)rowval, counts, index_base, d_lib.p, scale_factor, mean, var);
        else
            upload_moments_pipeline<int32_t>(a, colptr, (const int32_t *)rowval, counts, index_base, d_lib.p, scale_factor, mean, var);
        csc_validate(a, "svb_csc_upload_lognorm_moments");
    } catch (...) {
        delete a;
        throw;
    }
    *out = a;
    SVB_API_END
}

int svb_row_sums(svb_matrix_t a, int64_t *s) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && s, SVB_EARG, "svb_row_sums: null argument");
    SVB_CHECK(a->vtype == SVB_I32, SVB_EARG, "svb_row_sums: integer counts required (normalize.jl:24)");
    cudaStream_t st = ctx().stream;
    DevBuf<long long> d((size_t)std::max<int64_t>(a->nrow, 1));
    SVB_CUDA(cudaMemsetAsync(d.p, 0, (size_t)a->nrow * 8, st));
    if (a->nnz > 0) {
        row_sums_kernel<<<grid1(a->nnz), 256, 0, st>>>(a->rowidx, (const int32_t *)a->val, a->nnz, (unsigned long long *)d.p);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    SVB_CUDA(cudaMemcpyAsync(s, d.p, (size_t)a->nrow * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_API_END
}

static int normalize_impl(svb_matrix_t a, const int64_t *h_libsize, int method, double scale_factor, int dtype, svb_matrix_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && out, SVB_EARG, "svb_normalize: null argument");
    SVB_CHECK(a->vtype == SVB_I32, SVB_EARG, "svb_normalize: integer counts required");
    SVB_CHECK(method == SVB_NORM_LOGNORMALIZE || method == SVB_NORM_RELATIVECOUNTS, SVB_EARG, "unknown normalization method");
    SVB_CHECK(dtype == SVB_F32 || dtype == SVB_F64, SVB_EARG, "svb_normalize: dtype must be F32 or F64");
    cudaStream_t st = ctx().stream;
    svb_matrix_s *b = matrix_alloc(a->nrow, a->ncol, a->nnz, dtype);
    try {
        SVB_CUDA(cudaMemcpyAsync(b->colptr, a->colptr, (size_t)(a->ncol + 1) * 8, cudaMemcpyDeviceToDevice, st));
        if (a->nnz > 0) {
            SVB_CUDA(cudaMemcpyAsync(b->rowidx, a->rowidx, (size_t)a->nnz * 4, cudaMemcpyDeviceToDevice, st));
            DevBuf<long long> s((size_t)std::max<int64_t>(a->nrow, 1));
            if (h_libsize) {
                SVB_CUDA(cudaMemcpyAsync(s.p, h_libsize, (size_t)a->nrow * 8, cudaMemcpyHostToDevice, st));
            } else {
                SVB_CUDA(cudaMemsetAsync(s.p, 0, (size_t)a->nrow * 8, st));
                row_sums_kernel<<<grid1(a->nnz), 256, 0, st>>>(a->rowidx, (const int32_t *)a->val, a->nnz, (unsigned long long *)s.p);
                count_launch();
            }
            const int do_log = method == SVB_NORM_LOGNORMALIZE;
            if (dtype == SVB_F64)
                libnorm_kernel<double><<<grid1(a->nnz), 256, 0, st>>>(a->rowidx, (const int32_t *)a->val, a->nnz, s.p, scale_factor, do_log, (double *)b->val);
            else
                libnorm_kernel<float><<<grid1(a->nnz), 256, 0, st>>>(a->rowidx, (const int32_t *)a->val, a->nnz, s.p, (float)scale_factor, do_log, (float *)b->val);
            count_launch();
            SVB_LAUNCH_CHECK();
            SVB_CUDA(cudaStreamSynchronize(st));
        }
        SVB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        delete b;
        throw;
    }
    *out = b;
    SVB_API_END
}

int svb_normalize(svb_matrix_t a, int method, double scale_factor, int dtype, svb_matrix_t *out) {
    return normalize_impl(a, nullptr, method, scale_factor, dtype, out);
}

int svb_normalize_libsize(svb_matrix_t a, const int64_t *libsize, int method, double scale_factor, int dtype, svb_matrix_t *out) {
    if (!libsize) {
        svb::set_last_error("svb_normalize_libsize: null library sizes");
        return SVB_EARG;
    }
    return normalize_impl(a, libsize, method, scale_factor, dtype, out);
}

int svb_mean_var(svb_matrix_t a, double *mu, double *var) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && mu && var, SVB_EARG, "svb_mean_var: null argument");
    DevBuf<double> d_mu((size_t)std::max<int64_t>(a->ncol, 1)), d_var((size_t)std::max<int64_t>(a->ncol, 1));
    mean_var_device(a, d_mu.p, d_var.p);
    cudaStream_t st = ctx().stream;
    SVB_CUDA(cudaMemcpyAsync(mu, d_mu.p, (size_t)a->ncol * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaMemcpyAsync(var, d_var.p, (size_t)a->ncol * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_API_END
}

int svb_welford_carry(svb_matrix_t a, int64_t *count, double *mu, double *s) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && count && mu && s, SVB_EARG, "svb_welford_carry: null argument");
    SVB_CHECK(a->vtype == SVB_I32 || a->vtype == SVB_F64, SVB_EARG, "svb_welford_carry: Int or Float64 data (Float64 chain)");
    if (a->ncol == 0) return SVB_OK;
    cudaStream_t st = ctx().stream;
    const size_t nc = (size_t)a->ncol;
    DevBuf<long long> d_c(nc);
    DevBuf<double> d_mu(nc), d_s(nc);
    SVB_CUDA(cudaMemcpyAsync(d_c.p, count, nc * 8, cudaMemcpyHostToDevice, st));
    SVB_CUDA(cudaMemcpyAsync(d_mu.p, mu, nc * 8, cudaMemcpyHostToDevice, st));
    SVB_CUDA(cudaMemcpyAsync(d_s.p, s, nc * 8, cudaMemcpyHostToDevice, st));
    DevBuf<int32_t> order;
    column_order(a, order);
    const unsigned grid = (unsigned)((a->ncol + 63) / 64);
    if (a->vtype == SVB_I32)
        welford_carry_kernel<int32_t><<<grid, 64, 0, st>>>(a->colptr, (const int32_t *)a->val, order.p, a->ncol, d_c.p, d_mu.p, d_s.p);
    else
        welford_carry_kernel<double><<<grid, 64, 0, st>>>(a->colptr, (const double *)a->val, order.p, a->ncol, d_c.p, d_mu.p, d_s.p);
    count_launch();
    SVB_LAUNCH_CHECK();
    SVB_CUDA(cudaMemcpyAsync(count, d_c.p, nc * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaMemcpyAsync(mu, d_mu.p, nc * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaMemcpyAsync(s, d_s.p, nc * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_API_END
}

int svb_stdvar_clipped(svb_matrix_t a, const double *mu, const double *sd, double vmax, double *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && mu && sd && out, SVB_EARG, "svb_stdvar_clipped: null argument");
    SVB_CHECK(a->vtype == SVB_I32, SVB_EARG, "svb_stdvar_clipped: integer counts required (variablefeatures.jl:21)");
    if (!(vmax > 0.0)) vmax = std::sqrt((double)a->nrow);
    cudaStream_t st = ctx().stream;
    const size_t nc = (size_t)std::max<int64_t>(a->ncol, 1);
    DevBuf<double> d_mu(nc), d_sd(nc), d_out(nc);
    SVB_CUDA(cudaMemcpyAsync(d_mu.p, mu, (size_t)a->ncol * 8, cudaMemcpyHostToDevice, st));
    SVB_CUDA(cudaMemcpyAsync(d_sd.p, sd, (size_t)a->ncol * 8, cudaMemcpyHostToDevice, st));
    if (a->ncol > 0) {
        DevBuf<int32_t> order;
        column_order(a, order);
        const unsigned grid = (unsigned)((a->ncol * 32 + 255) / 256);
        stdvar_kernel<<<grid, 256, 0, st>>>(a->colptr, (const int32_t *)a->val, order.p, a->ncol, a->nrow, d_mu.p, d_sd.p, vmax, d_out.p);
        count_launch();
        SVB_LAUNCH_CHECK();
        SVB_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)a->ncol * 8, cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
    }
    SVB_API_END
}

static int scale_impl(svb_matrix_t a, const double *h_mean, const double *h_var, double scale_max, int dtype,
                      svb_matrix_t *out, double *mu_out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && out && mu_out, SVB_EARG, "svb_scale: null argument");
    SVB_CHECK(dtype == SVB_F32 || dtype == SVB_F64, SVB_EARG, "svb_scale: dtype must be F32 or F64");
    SVB_CHECK(!(a->vtype == SVB_F32 && dtype == SVB_F64), SVB_EARG, "svb_scale: Float32 data scales to Float32");
    cudaStream_t st = ctx().stream;
    const size_t nc = (size_t)std::max<int64_t>(a->ncol, 1);
    DevBuf<double> d_mean(nc), d_var(nc), d_sd(nc), d_mus(nc), d_smax(nc);
    if (h_mean) {
        SVB_CUDA(cudaMemcpyAsync(d_mean.p, h_mean, (size_t)a->ncol * 8, cudaMemcpyHostToDevice, st));
        SVB_CUDA(cudaMemcpyAsync(d_var.p, h_var, (size_t)a->ncol * 8, cudaMemcpyHostToDevice, st));
    } else {
        mean_var_device(a, d_mean.p, d_var.p);
    }
    svb_matrix_s *b = matrix_alloc(a->nrow, a->ncol, a->nnz, dtype);
    try {
        SVB_CUDA(cudaMemcpyAsync(b->colptr, a->colptr, (size_t)(a->ncol + 1) * 8, cudaMemcpyDeviceToDevice, st));
        if (a->ncol > 0) {
            const unsigned g = (unsigned)((a->ncol + 255) / 256);
            if (dtype == SVB_F64) scale_prepare_kernel<double><<<g, 256, 0, st>>>(d_mean.p, d_var.p, a->ncol, scale_max, d_sd.p, d_mus.p, d_smax.p);
            else scale_prepare_kernel<float><<<g, 256, 0, st>>>(d_mean.p, d_var.p, a->ncol, scale_max, d_sd.p, d_mus.p, d_smax.p);
            count_launch();
        }
        if (a->nnz > 0) {
            SVB_CUDA(cudaMemcpyAsync(b->rowidx, a->rowidx, (size_t)a->nnz * 4, cudaMemcpyDeviceToDevice, st));
            dim3 grid((unsigned)a->ncol, 8);
            if (a->vtype == SVB_F64 && dtype == SVB_F64)
                scale_apply_kernel<double, double><<<grid, 256, 0, st>>>(a->colptr, (const double *)a->val, d_sd.p, d_smax.p, (double *)b->val);
            else if (a->vtype == SVB_F64 && dtype == SVB_F32)
                scale_apply_kernel<double, float><<<grid, 256, 0, st>>>(a->colptr, (const double *)a->val, d_sd.p, d_smax.p, (float *)b->val);
            else if (a->vtype == SVB_I32 && dtype == SVB_F64)
                scale_apply_kernel<int32_t, double><<<grid, 256, 0, st>>>(a->colptr, (const int32_t *)a->val, d_sd.p, d_smax.p, (double *)b->val);
            else if (a->vtype == SVB_I32 && dtype == SVB_F32)
                scale_apply_kernel<int32_t, float><<<grid, 256, 0, st>>>(a->colptr, (const int32_t *)a->val, d_sd.p, d_smax.p, (float *)b->val);
            else
                scale_apply_f32_kernel<<<grid, 256, 0, st>>>(a->colptr, (const float *)a->val, d_sd.p, d_smax.p, (float *)b->val);
            count_launch();
            SVB_LAUNCH_CHECK();
        }
        SVB_CUDA(cudaMemcpyAsync(mu_out, d_mus.p, (size_t)a->ncol * 8, cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        delete b;
        throw;
    }
    *out = b;
    SVB_API_END
}

int svb_scale(svb_matrix_t a, double scale_max, int dtype, svb_matrix_t *out, double *mu_out) {
    return scale_impl(a, nullptr, nullptr, scale_max, dtype, out, mu_out);
}

int svb_scale_with_moments(svb_matrix_t a, const double *mean, const double *var, double scale_max, int dtype,
                           svb_matrix_t *out, double *mu_out) {
    if (!mean || !var) {
        svb::set_last_error("svb_scale_with_moments: null moments");
        return SVB_EARG;
    }
    return scale_impl(a, mean, var, scale_max, dtype, out, mu_out);
}

}  // extern "C"

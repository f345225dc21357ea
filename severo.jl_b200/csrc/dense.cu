// dense.cu — tall-skinny fp64 kernels of the Lanczos bidiagonalisation (the BLAS-2/3 part of
// libcell's irlba, call site src/irlba.jl:66-71): classical Gram-Schmidt reorthogonalisation
// (t = X'y ; y -= X t), norms / normalisation, and the restart / final products W*P, V*Q.
// All of them are HBM-bound streams over the m x w basis; reductions are two-stage with a
// last-block pass in a fixed order, so every result is run-to-run deterministic.
#include "svb_internal.h"
#include "p2p.cuh"
#include <cooperative_groups.h>

#include <algorithm>

namespace svb {

// ---- persistent scratch for two-stage reductions ----------------------------------------------
struct Scratch {
    double *partials = nullptr;
    size_t cap = 0;
    unsigned int *counters = nullptr;
    size_t ncounters = 0;
};
static Scratch &scr_of_ctx() {
    Context &C = ctx();
    if (!C.dense_scratch) C.dense_scratch = new Scratch();
    return *static_cast<Scratch *>(C.dense_scratch);
}
#define g_scr (scr_of_ctx())

static void scratch_reserve(size_t ndoubles, size_t ncounters) {
    cudaStream_t st = ctx().stream;
    if (ndoubles > g_scr.cap) {
        SVB_CUDA(cudaStreamSynchronize(st));
        if (g_scr.partials) cudaFree(g_scr.partials);
        g_scr.cap = std::max<size_t>(ndoubles, 1u << 16);
        SVB_CUDA(cudaMalloc((void **)&g_scr.partials, g_scr.cap * sizeof(double)));
    }
    if (ncounters > g_scr.ncounters) {
        SVB_CUDA(cudaStreamSynchronize(st));
        if (g_scr.counters) cudaFree(g_scr.counters);
        g_scr.ncounters = std::max<size_t>(ncounters, 4096);
        SVB_CUDA(cudaMalloc((void **)&g_scr.counters, g_scr.ncounters * sizeof(unsigned int)));
        SVB_CUDA(cudaMemsetAsync(g_scr.counters, 0, g_scr.ncounters * sizeof(unsigned int), st));
    }
}

constexpr int CT = 8;        // columns per CTA in gemv_t
constexpr int TS_THREADS = 256;
constexpr int TS_RPT = 4;    // rows per thread per sweep

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// t[c] = sum_i X[i + c*ld] * y[i].  grid (nrb, ceil(j/CT)).
__global__ void __launch_bounds__(TS_THREADS) ts_gemv_t_kernel(const double *__restrict__ X, int64_t ld, int64_t L, int j,
                                                               const double *__restrict__ y, double *__restrict__ t,
                                                               double *__restrict__ partials, unsigned int *counters,
                                                               int use_p2p, P2PCtx pc) {
    __shared__ double sh[TS_THREADS / 32][CT];
    __shared__ bool is_last;
    const int c0 = blockIdx.y * CT;
    const int nc = min(CT, j - c0);
    double acc[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[c] = 0.0;
    const int64_t chunk = (int64_t)TS_THREADS * TS_RPT;
    for (int64_t base = (int64_t)blockIdx.x * chunk; base < L; base += (int64_t)gridDim.x * chunk) {
#pragma unroll
        for (int r = 0; r < TS_RPT; ++r) {
            const int64_t i = base + (int64_t)r * TS_THREADS + threadIdx.x;
            if (i < L) {
                const double yv = y[i];
                const double *xp = X + i + (int64_t)c0 * ld;
#pragma unroll
                for (int c = 0; c < CT; ++c)
                    if (c < nc) acc[c] = fma(__ldg(xp + (int64_t)c * ld), yv, acc[c]);
            }
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const double s = warp_sum(acc[c]);
        if (lane == 0) sh[wid][c] = s;
    }
    __syncthreads();
    double *mypart = partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * CT;
    if (threadIdx.x < CT) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < TS_THREADS / 32; ++w) s += sh[w][threadIdx.x];
        mypart[threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(&counters[blockIdx.y], 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // warp `wid` sums column `wid` over all row blocks in a fixed order
        if (wid < nc) {
            const double *p = partials + (size_t)blockIdx.y * gridDim.x * CT + wid;
            double s = 0.0;
            for (unsigned int b = lane; b < gridDim.x; b += 32) s += __ldcg(p + (size_t)b * CT);
            s = warp_sum(s);
            if (lane == 0) {
                if (use_p2p) p2p_store(pc, c0 + wid, s);  // fused exchange: coefficients go straight to every rank
                else t[c0 + wid] = s;
            }
        }
        if (threadIdx.x == 0) counters[blockIdx.y] = 0u;
        if (use_p2p) p2p_publish_last_block(pc, gridDim.y);  // the last column tile to finish publishes the epoch
    }
}

// y[i] = beta*y[i] + alpha * sum_c X[i + c*ld] * t[c] ; optional nrm2_out = sum_i y[i]^2
__global__ void __launch_bounds__(TS_THREADS) ts_gemv_n_kernel(const double *__restrict__ X, int64_t ld, int64_t L, int j,
                                                               const double *__restrict__ t, double alpha, double beta,
                                                               double *__restrict__ y, double *__restrict__ nrm2_out,
                                                               double *__restrict__ partials, unsigned int *counter,
                                                               int t_p2p, P2PCtx pt, int nrm_p2p, P2PCtx pn) {
    __shared__ double sh[TS_THREADS / 32];
    __shared__ double ts[TS_THREADS];
    __shared__ bool is_last;
    const bool smem_t = j <= TS_THREADS;
    if (t_p2p) p2p_wait(pt);  // fused exchange: every rank's coefficients are in the local mailbox
    if (smem_t) {
        if ((int)threadIdx.x < j) ts[threadIdx.x] = t_p2p ? p2p_sum(pt, threadIdx.x) : t[threadIdx.x];
        __syncthreads();
    }
    double ss = 0.0;
    const int64_t chunk = (int64_t)TS_THREADS * 2;
    for (int64_t base = (int64_t)blockIdx.x * chunk; base < L; base += (int64_t)gridDim.x * chunk) {
        const int64_t i0 = base + threadIdx.x, i1 = i0 + TS_THREADS;
        const bool v0 = i0 < L, v1 = i1 < L;
        double a0 = 0.0, a1 = 0.0;
        int c = 0;
        for (; c + 4 <= j; c += 4) {
            const double t0 = smem_t ? ts[c] : __ldg(t + c), t1 = smem_t ? ts[c + 1] : __ldg(t + c + 1);
            const double t2 = smem_t ? ts[c + 2] : __ldg(t + c + 2), t3 = smem_t ? ts[c + 3] : __ldg(t + c + 3);
            const double *xp = X + (int64_t)c * ld;
            if (v0) {
                const double x0 = __ldg(xp + i0), x1 = __ldg(xp + ld + i0), x2 = __ldg(xp + 2 * ld + i0), x3 = __ldg(xp + 3 * ld + i0);
                a0 = fma(x0, t0, a0); a0 = fma(x1, t1, a0); a0 = fma(x2, t2, a0); a0 = fma(x3, t3, a0);
            }
            if (v1) {
                const double x0 = __ldg(xp + i1), x1 = __ldg(xp + ld + i1), x2 = __ldg(xp + 2 * ld + i1), x3 = __ldg(xp + 3 * ld + i1);
                a1 = fma(x0, t0, a1); a1 = fma(x1, t1, a1); a1 = fma(x2, t2, a1); a1 = fma(x3, t3, a1);
            }
        }
        for (; c < j; ++c) {
            const double tc = smem_t ? ts[c] : __ldg(t + c);
            if (v0) a0 = fma(__ldg(X + (int64_t)c * ld + i0), tc, a0);
            if (v1) a1 = fma(__ldg(X + (int64_t)c * ld + i1), tc, a1);
        }
        if (v0) {
            double r = alpha * a0;
            if (beta != 0.0) r = fma(beta, y[i0], r);
            y[i0] = r;
            ss = fma(r, r, ss);
        }
        if (v1) {
            double r = alpha * a1;
            if (beta != 0.0) r = fma(beta, y[i1], r);
            y[i1] = r;
            ss = fma(r, r, ss);
        }
    }
    if (nrm2_out == nullptr) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ss = warp_sum(ss);
    if (lane == 0) sh[wid] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < TS_THREADS / 32; ++w) s += sh[w];
        partials[blockIdx.x] = s;
        __threadfence();
        const unsigned int ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        if (wid == 0) {
            __threadfence();
            double s = 0.0;
            for (unsigned int b = lane; b < gridDim.x; b += 32) s += __ldcg(partials + b);
            s = warp_sum(s);
            if (lane == 0) {
                if (nrm_p2p) p2p_store(pn, 0, s);  // fused exchange of the local |y|^2
                else *nrm2_out = s;
                *counter = 0u;
            }
        }
        if (nrm_p2p) p2p_publish_last_block(pn, 1);
    }
}

static int row_blocks(int64_t L, int64_t rows_per_cta, int max_per_sm) {
    const int64_t need = (L + rows_per_cta - 1) / rows_per_cta;
    return (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)ctx().sm_count * max_per_sm));
}

// CTAs of `kernel` that are resident at once on the whole chip (one full wave); the streaming kernels use
// grid-stride loops, so a grid of exactly this size avoids the partial second wave ncu showed (1.33-1.67 waves).
template <typename K>
static int wave_slots(K kernel, int threads) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    return per_sm * ctx().sm_count;
}

void ts_gemv_t(const double *X, int64_t ld, int64_t L, int j, const double *y, double *t, int cls, const P2PCtx *produce_t) {
    if (j <= 0) return;
    const int ctiles = (j + CT - 1) / CT;
    // keep the total CTA count near 4 waves regardless of the number of column tiles
    static int slots_t = 0;
    if (!slots_t) slots_t = wave_slots(ts_gemv_t_kernel, TS_THREADS);
    const int64_t need_t = (L + TS_THREADS * TS_RPT - 1) / (TS_THREADS * TS_RPT);
    int nrb = (int)std::max<int64_t>(1, std::min<int64_t>(need_t, std::max(1, slots_t / ctiles)));
    scratch_reserve((size_t)ctiles * nrb * CT + 4096, (size_t)ctiles + 8);
    KTimer kt(cls, 8.0 * ((double)L * j + (double)L * ((j + CT - 1) / CT)));
    dim3 grid((unsigned)nrb, (unsigned)ctiles);
    ts_gemv_t_kernel<<<grid, TS_THREADS, 0, ctx().stream>>>(X, ld, L, j, y, t, g_scr.partials + 4096, g_scr.counters + 8,
                                                            produce_t ? 1 : 0, produce_t ? *produce_t : P2PCtx{});
    SVB_LAUNCH_CHECK();
}

void ts_gemv_n(const double *X, int64_t ld, int64_t L, int j, const double *t, double alpha, double beta, double *y,
               double *nrm2_out, int cls, const P2PCtx *consume_t, const P2PCtx *produce_nrm) {
    static int slots_n = 0;
    if (!slots_n) slots_n = wave_slots(ts_gemv_n_kernel, TS_THREADS);
    const int nrb = (int)std::max<int64_t>(1, std::min<int64_t>((L + TS_THREADS * 2 - 1) / (TS_THREADS * 2), slots_n));
    scratch_reserve(4096 + 64, 8);
    KTimer kt(cls, 8.0 * ((double)L * j + 2.0 * L));
    ts_gemv_n_kernel<<<(unsigned)nrb, TS_THREADS, 0, ctx().stream>>>(X, ld, L, j, t, alpha, beta, y, nrm2_out, g_scr.partials,
                                                                     g_scr.counters, consume_t ? 1 : 0,
                                                                     consume_t ? *consume_t : P2PCtx{}, produce_nrm ? 1 : 0,
                                                                     produce_nrm ? *produce_nrm : P2PCtx{});
    SVB_LAUNCH_CHECK();
}

// ---- restart / final products: out[:, 0..k) = X[:, 0..w) * P[0..w, 0..k) (* colscale) --------------
constexpr int GM_THREADS = 128;
constexpr int GM_CC = 8;

__global__ void __launch_bounds__(GM_THREADS) ts_gemm_kernel(const double *__restrict__ X, int64_t ld, int64_t L, int w,
                                                             const double *__restrict__ P, int ldp, int k, int kp,
                                                             double *__restrict__ out, int64_t ldo,
                                                             const double *__restrict__ colscale) {
    extern __shared__ double Ps[];  // [w][kp], zero padded
    for (int idx = threadIdx.x; idx < w * kp; idx += blockDim.x) {
        const int l = idx / kp, c = idx - l * kp;
        double v = (c < k) ? P[l + (int64_t)c * ldp] : 0.0;
        if (colscale && c < k) v *= colscale[c];
        Ps[idx] = v;
    }
    __syncthreads();
    const int64_t i0 = (int64_t)blockIdx.x * (GM_THREADS * 2) + threadIdx.x;
    const int64_t i1 = i0 + GM_THREADS;
    const bool v0 = i0 < L, v1 = i1 < L;
    const int64_t j0 = v0 ? i0 : 0, j1 = v1 ? i1 : 0;
    for (int c0 = 0; c0 < kp; c0 += GM_CC) {
        double a0[GM_CC], a1[GM_CC];
#pragma unroll
        for (int c = 0; c < GM_CC; ++c) { a0[c] = 0.0; a1[c] = 0.0; }
#pragma unroll 2
        for (int l = 0; l < w; ++l) {
            const double x0 = __ldg(X + (int64_t)l * ld + j0);
            const double x1 = __ldg(X + (int64_t)l * ld + j1);
            const double2 *pp = reinterpret_cast<const double2 *>(Ps + l * kp + c0);
#pragma unroll
            for (int c = 0; c < GM_CC / 2; ++c) {
                const double2 p = pp[c];
                a0[2 * c] = fma(x0, p.x, a0[2 * c]);
                a0[2 * c + 1] = fma(x0, p.y, a0[2 * c + 1]);
                a1[2 * c] = fma(x1, p.x, a1[2 * c]);
                a1[2 * c + 1] = fma(x1, p.y, a1[2 * c + 1]);
            }
        }
#pragma unroll
        for (int c = 0; c < GM_CC; ++c) {
            if (c0 + c < k) {
                if (v0) out[i0 + (int64_t)(c0 + c) * ldo] = a0[c];
                if (v1) out[i1 + (int64_t)(c0 + c) * ldo] = a1[c];
            }
        }
    }
}

// fp64 tensor-core version (mma.sync.m8n8k4 DMMA; tcgen05 has no fp64): one warp owns 32 rows (4 m-tiles),
// walks the padded K = w in steps of 4 and carries DM_NT = 4 n-tiles (32 output columns) of accumulators per
// pass. A fragments come straight from the column-major basis (8 consecutive rows of 4 columns per load),
// B fragments from the zero-padded P staged in shared memory. Fragment layout (PTX ISA, m8n8k4 .f64):
// a0 = A[lane>>2][lane&3], b0 = B[lane&3][lane>>2], c0/c1 = C[lane>>2][2*(lane&3) + {0,1}].
constexpr int DM_THREADS = 128;
constexpr int DM_NT = 4;  // n-tiles per pass (7 = one pass for nu = 50 was slower: 160 registers)

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(DM_THREADS) ts_gemm_dmma_kernel(const double *__restrict__ X, int64_t ld, int64_t L, int w,
                                                                  const double *__restrict__ P, int ldp, int k, int kp8, int wp4,
                                                                  double *__restrict__ out, int64_t ldo,
                                                                  const double *__restrict__ colscale) {
    extern __shared__ double Ps[];  // [wp4][kp8], zero padded
    for (int idx = threadIdx.x; idx < wp4 * kp8; idx += blockDim.x) {
        const int l = idx / kp8, c = idx - l * kp8;
        double v = (l < w && c < k) ? P[l + (int64_t)c * ldp] : 0.0;
        if (colscale && c < k) v *= colscale[c];
        Ps[idx] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t r0 = ((int64_t)blockIdx.x * (DM_THREADS / 32) + warp) * 32;
    if (r0 >= L) return;
    for (int n0 = 0; n0 < kp8; n0 += 8 * DM_NT) {
        double acc[4][DM_NT][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < DM_NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        for (int l0 = 0; l0 < wp4; l0 += 4) {
            const int l = l0 + t;
            double a[4], b[DM_NT];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                const int64_t row = r0 + mt * 8 + g;
                a[mt] = (l < w && row < L) ? __ldg(X + (int64_t)l * ld + row) : 0.0;
            }
#pragma unroll
            for (int nt = 0; nt < DM_NT; ++nt) {
                const int col = n0 + nt * 8 + g;
                b[nt] = (col < kp8) ? Ps[l * kp8 + col] : 0.0;
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < DM_NT; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int64_t row = r0 + mt * 8 + g;
            if (row < L) {
#pragma unroll
                for (int nt = 0; nt < DM_NT; ++nt) {
                    const int col = n0 + nt * 8 + 2 * t;
                    if (col < k) out[row + (int64_t)col * ldo] = acc[mt][nt][0];
                    if (col + 1 < k) out[row + (int64_t)(col + 1) * ldo] = acc[mt][nt][1];
                }
            }
        }
    }
}

// Round 2 (C4 regime: w = 107, k ~ 100 — the product is bound by the fp64 pipe, 2*L*w*k flop): the first DMMA kernel walked K
// once per group of 4 n-tiles, re-reading its A fragments from L1/L2 up to 4 times, and every 128-row CTA staged the whole
// of P again (90 KB at C4, 3 x the bytes of its X tile): 0.92 TB/s. Here the CTAs are PERSISTENT (P is staged once per CTA and
// reused for every row block) and a row block of 32 rows is shared by four warps that split the n-tiles among them, so every
// warp walks K exactly once with all its accumulators live (NT n-tiles x 4 m-tiles x 2 doubles) and the four warps' identical
// A loads hit L1. 8 warps per CTA = 2 row groups x 4 column groups.
constexpr int DP_THREADS = 256;
template <int NT, int CG>
__global__ void __launch_bounds__(DP_THREADS, (NT <= 4 ? 2 : 1)) ts_gemm_dmma_persistent_kernel(const double *__restrict__ X, int64_t ld, int64_t L, int w,
                                                                             const double *__restrict__ P, int ldp, int k, int kp8,
                                                                             int wp4, double *__restrict__ out, int64_t ldo,
                                                                             const double *__restrict__ colscale) {
    extern __shared__ double Ps[];  // [wp4][kp8], zero padded
    for (int idx = threadIdx.x; idx < wp4 * kp8; idx += blockDim.x) {
        const int l = idx / kp8, c = idx - l * kp8;
        double v = (l < w && c < k) ? P[l + (int64_t)c * ldp] : 0.0;
        if (colscale && c < k) v *= colscale[c];
        Ps[idx] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    // CG column groups share a row block of 32 rows (their identical A loads hit L1), 8 / CG row groups per CTA
    constexpr int RG = 8 / CG, ROWS = RG * 32;
    const int rg = warp / CG, cg = warp % CG;
    const int ntiles = kp8 >> 3;
    const int64_t nblk = (L + ROWS - 1) / ROWS;
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int64_t r0 = blk * ROWS + rg * 32;
        if (r0 >= L) continue;
        double acc[4][NT][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int i = 0; i < NT; ++i) acc[mt][i][0] = acc[mt][i][1] = 0.0;
        const double *xrow[4];
        bool rok[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int64_t row = r0 + mt * 8 + g;
            rok[mt] = row < L;
            xrow[mt] = X + (rok[mt] ? row : 0);
        }
        // five k-steps in flight: 20 loads of 256 B per warp (with two the restart product ran at 1.9 TB/s — too few bytes in
        // flight per SM for HBM)
#pragma unroll 5
        for (int l0 = 0; l0 < wp4; l0 += 4) {
            const int l = l0 + t;
            const bool lok = l < w;
            double a[4], b[NT];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) a[mt] = (lok && rok[mt]) ? __ldg(xrow[mt] + (int64_t)l * ld) : 0.0;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const int nt = cg + CG * i;
                b[i] = (nt < ntiles) ? Ps[l * kp8 + nt * 8 + g] : 0.0;
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int i = 0; i < NT; ++i) dmma_m8n8k4(acc[mt][i][0], acc[mt][i][1], a[mt], b[i]);
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int64_t row = r0 + mt * 8 + g;
            if (row < L) {
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    const int col = (cg + CG * i) * 8 + 2 * t;
                    if (col < k) out[row + (int64_t)col * ldo] = acc[mt][i][0];
                    if (col + 1 < k) out[row + (int64_t)(col + 1) * ldo] = acc[mt][i][1];
                }
            }
        }
    }
}

template <int NT, int CG>
static void launch_dmma_persistent(const double *X, int64_t ld, int64_t L, int w, const double *P, int ldp, int k, int kp8, int wp4,
                                   double *out, int64_t ldo, const double *colscale_dev, size_t smem) {
    auto kern = ts_gemm_dmma_persistent_kernel<NT, CG>;
    if (smem > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DP_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int64_t nblk = (L + (8 / CG) * 32 - 1) / ((8 / CG) * 32);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nblk, (int64_t)per_sm * ctx().sm_count));
    kern<<<grid, DP_THREADS, smem, ctx().stream>>>(X, ld, L, w, P, ldp, k, kp8, wp4, out, ldo, colscale_dev);
}

void ts_gemm(const double *X, int64_t ld, int64_t L, int w, const double *P, int ldp, int k, double *out, int64_t ldo,
             const double *colscale_dev) {
    if (k <= 0 || L <= 0) return;
    static const bool use_dmma = getenv("SVB_NO_DMMA") == nullptr;
    if (use_dmma) {
        const int kp8 = (k + 7) / 8 * 8, wp4 = (w + 3) / 4 * 4;
        const size_t smem = (size_t)wp4 * kp8 * sizeof(double);
        static const bool persistent = !(getenv("SVB_DMMA_PERSISTENT") && atoi(getenv("SVB_DMMA_PERSISTENT")) == 0);
        if (smem <= ctx().smem_optin && persistent && kp8 <= 256) {
            const int ntl = kp8 / 8;
            static const int cg_env = getenv("SVB_DMMA_CG") ? atoi(getenv("SVB_DMMA_CG")) : 0;
            KTimer kt(SVB_K_RESTART, 8.0 * ((double)L * w + (double)L * k));
            // n-tiles per warp: two column groups (half the redundant A loads) while a warp's accumulators fit (<= 4 tiles)
            if (ntl <= 8 && cg_env != 4) {
                if (ntl <= 4) launch_dmma_persistent<2, 2>(X, ld, L, w, P, ldp, k, kp8, wp4, out, ldo, colscale_dev, smem);
                else launch_dmma_persistent<4, 2>(X, ld, L, w, P, ldp, k, kp8, wp4, out, ldo, colscale_dev, smem);
            } else {
                const int nt = (ntl + 3) / 4;  // four column groups
                if (nt <= 2) launch_dmma_persistent<2, 4>(X, ld, L, w, P, ldp, k, kp8, wp4, out, ldo, colscale_dev, smem);
                else if (nt <= 4) launch_dmma_persistent<4, 4>(X, ld, L, w, P, ldp, k, kp8, wp4, out, ldo, colscale_dev, smem);
                else launch_dmma_persistent<8, 4>(X, ld, L, w, P, ldp, k, kp8, wp4, out, ldo, colscale_dev, smem);
            }
            SVB_LAUNCH_CHECK();
            return;
        }
        if (smem <= ctx().smem_optin) {
            if (smem > 48 * 1024)
                SVB_CUDA(cudaFuncSetAttribute(ts_gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int64_t rows_per_cta = (DM_THREADS / 32) * 32;
            const int64_t grid = (L + rows_per_cta - 1) / rows_per_cta;
            KTimer kt(SVB_K_RESTART, 8.0 * ((double)L * w + (double)L * k));
            ts_gemm_dmma_kernel<<<(unsigned)grid, DM_THREADS, smem, ctx().stream>>>(X, ld, L, w, P, ldp, k, kp8, wp4, out, ldo,
                                                                                    colscale_dev);
            SVB_LAUNCH_CHECK();
            return;
        }
    }
    const int kp = (k + GM_CC - 1) / GM_CC * GM_CC;
    const size_t smem = (size_t)w * kp * sizeof(double);
    SVB_CHECK(smem <= ctx().smem_optin, SVB_EDIM, "restart product: work size too large for shared memory");
    if (smem > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(ts_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (L + GM_THREADS * 2 - 1) / (GM_THREADS * 2);
    KTimer kt(SVB_K_RESTART, 8.0 * ((double)L * w + (double)L * k));
    ts_gemm_kernel<<<(unsigned)grid, GM_THREADS, smem, ctx().stream>>>(X, ld, L, w, P, ldp, k, kp, out, ldo, colscale_dev);
    SVB_LAUNCH_CHECK();
}

// ---- vector kernels --------------------------------------------------------------------------------
__global__ void __launch_bounds__(TS_THREADS) sumsq_kernel(const double *__restrict__ x, int64_t L, double *out, double *partials,
                                                           unsigned int *counter) {
    __shared__ double sh[TS_THREADS / 32];
    __shared__ bool is_last;
    double ss = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        ss = fma(v, v, ss);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ss = warp_sum(ss);
    if (lane == 0) sh[wid] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < TS_THREADS / 32; ++w) s += sh[w];
        partials[blockIdx.x] = s;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && wid == 0) {
        __threadfence();
        double s = 0.0;
        for (unsigned int b = lane; b < gridDim.x; b += 32) s += __ldcg(partials + b);
        s = warp_sum(s);
        if (lane == 0) {
            *out = s;
            *counter = 0u;
        }
    }
}

void vec_sumsq(const double *x, int64_t L, double *out) {
    const int nrb = row_blocks(L, TS_THREADS * 4, 4);
    scratch_reserve(4096 + 64, 8);
    KTimer kt(SVB_K_VECTOR, 8.0 * L);
    sumsq_kernel<<<(unsigned)nrb, TS_THREADS, 0, ctx().stream>>>(x, L, out, g_scr.partials, g_scr.counters + 1);
    SVB_LAUNCH_CHECK();
}

__global__ void normalize_kernel(const double *__restrict__ x, int64_t L, const double *__restrict__ nrm2, double *__restrict__ y,
                                 double *norm_out, int *flag, double eps, int nrm_p2p, P2PCtx pn) {
    if (nrm_p2p) p2p_wait(pn);  // fused exchange: sum of the ranks' local |x|^2 (rank order, same bits everywhere)
    const double nrm = sqrt(nrm_p2p ? p2p_sum(pn, 0) : *nrm2);
    const double inv = 1.0 / nrm;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) y[i] = x[i] * inv;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (norm_out) *norm_out = nrm;
        if (flag && !(nrm >= eps)) *flag = 1;
    }
}

void vec_normalize(const double *x, int64_t L, const double *nrm2_dev, double *y, double *norm_out, int *flag_dev, double eps,
                   const P2PCtx *consume_nrm) {
    const int nrb = row_blocks(L, 256 * 4, 8);
    KTimer kt(SVB_K_VECTOR, 16.0 * L);
    normalize_kernel<<<(unsigned)nrb, 256, 0, ctx().stream>>>(x, L, nrm2_dev, y, norm_out, flag_dev, eps, consume_nrm ? 1 : 0,
                                                              consume_nrm ? *consume_nrm : P2PCtx{});
    SVB_LAUNCH_CHECK();
}

// ---- gene-side Gram-Schmidt step in ONE launch -------------------------------------------------------------------------
// The gene-side vectors (length n = 2,000 HVGs) live whole on every rank; one Lanczos step there was three launches of
// latency-bound kernels (coefficients V'f: 10 us, update f -= V h + |f|^2: 19 us on 4 CTAs, normalise: 5 us — ncu launch list of
// a C3 shard, profiles/r04_scaling.md), i.e. 34 us for 1.8 MB that sit in L2. Here a thread-block CLUSTER of 8 CTAs owns n/8
// rows each and exchanges its partial coefficients and partial |f|^2 through distributed shared memory: every CTA adds the
// eight partials in rank order (same bits in every CTA, run to run), two cluster barriers instead of two kernel boundaries.
namespace cg = cooperative_groups;
constexpr int VS_CL = 8, VS_T = 256, VS_MAXW = 256, VS_MAXROWS = 1024;

__global__ void __cluster_dims__(VS_CL, 1, 1) __launch_bounds__(VS_T)
vside_cgs_kernel(const double *__restrict__ V, int64_t n, int j, double *__restrict__ f, double *__restrict__ nrm2,
                 double *__restrict__ out, double *__restrict__ slot, int *__restrict__ flag, double eps) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), t = threadIdx.x;
    __shared__ double fs[VS_MAXROWS];
    __shared__ double ph[VS_CL][VS_MAXW];  // partial coefficients of every CTA of the cluster (written by their owners)
    __shared__ double h[VS_MAXW];
    __shared__ double red[4][64];
    __shared__ double pss[VS_CL];
    __shared__ double wred[VS_T / 32];
    const int64_t per = (n + VS_CL - 1) / VS_CL;
    const int64_t r0 = std::min<int64_t>(n, rank * per);
    const int nr = (int)(std::min<int64_t>(n, r0 + per) - r0);
    for (int i = t; i < nr; i += VS_T) fs[i] = f[r0 + i];
    cluster.sync();  // (also: every CTA of the cluster is running before anybody stores into its shared memory)
    // coefficients: thread = (column c of a group of 64, quarter q of this CTA's rows); a thread walks DOWN its column
    const int cl = t & 63, q = t >> 6;
    const int qrows = (nr + 3) / 4, qa = std::min(nr, q * qrows), qb = std::min(nr, qa + qrows);
    for (int c0 = 0; c0 < j; c0 += 64) {
        const int c = c0 + cl;
        double a0 = 0.0, a1 = 0.0;
        if (c < j) {
            const double *col = V + (int64_t)c * n + r0;
            int i = qa;
            for (; i + 1 < qb; i += 2) {
                a0 = fma(col[i], fs[i], a0);
                a1 = fma(col[i + 1], fs[i + 1], a1);
            }
            if (i < qb) a0 = fma(col[i], fs[i], a0);
        }
        red[q][cl] = a0 + a1;
        __syncthreads();
        if (q == 0 && c < j) {
            const double sum = (red[0][cl] + red[1][cl]) + (red[2][cl] + red[3][cl]);
            for (int p = 0; p < VS_CL; ++p) *cluster.map_shared_rank(&ph[rank][c], p) = sum;
        }
        __syncthreads();
    }
    cluster.sync();
    for (int c = t; c < j; c += VS_T) {
        double sum = 0.0;
#pragma unroll
        for (int p = 0; p < VS_CL; ++p) sum += ph[p][c];
        h[c] = sum;
    }
    __syncthreads();
    // update of this CTA's rows, |f|^2
    double ss = 0.0;
    for (int i = t; i < nr; i += VS_T) {
        const double *row = V + r0 + i;
        double a0 = 0.0, a1 = 0.0;
        int c = 0;
        for (; c + 1 < j; c += 2) {
            a0 = fma(row[(int64_t)c * n], h[c], a0);
            a1 = fma(row[(int64_t)(c + 1) * n], h[c + 1], a1);
        }
        if (c < j) a0 = fma(row[(int64_t)c * n], h[c], a0);
        const double v = fs[i] - (a0 + a1);
        fs[i] = v;
        ss = fma(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((t & 31) == 0) wred[t >> 5] = ss;
    __syncthreads();
    if (t == 0) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < VS_T / 32; ++k) sum += wred[k];
        for (int p = 0; p < VS_CL; ++p) *cluster.map_shared_rank(&pss[rank], p) = sum;
    }
    cluster.sync();
    double tot = 0.0;
#pragma unroll
    for (int p = 0; p < VS_CL; ++p) tot += pss[p];
    const double nrm = sqrt(tot), inv = 1.0 / nrm;
    for (int i = t; i < nr; i += VS_T) {
        f[r0 + i] = fs[i];
        if (out) out[r0 + i] = fs[i] * inv;
    }
    if (rank == 0 && t == 0) {
        *nrm2 = tot;
        if (out) {
            if (slot) *slot = nrm;
            if (flag && !(nrm >= eps)) *flag = 1;
        }
    }
}

bool vside_cgs_supported(int64_t n, int j) {
    static const bool on = !(getenv("SVB_VSIDE_FUSED") && atoi(getenv("SVB_VSIDE_FUSED")) == 0);
    return on && j >= 1 && j <= VS_MAXW && n >= 1 && n <= (int64_t)VS_CL * VS_MAXROWS;
}

// f <- f - V[:, :j] (V[:, :j]' f); *nrm2 = |f|^2; out (if given) = f/|f|, *slot = |f|, *flag = 1 when |f| < eps
void vside_cgs(const double *V, int64_t n, int j, double *f, double *nrm2, double *out, double *slot, int *flag, double eps) {
    KTimer kt(SVB_K_REORTH, 8.0 * (2.0 * (double)n * j + 4.0 * n));
    vside_cgs_kernel<<<VS_CL, VS_T, 0, ctx().stream>>>(V, n, j, f, nrm2, out, slot, flag, eps);
    SVB_LAUNCH_CHECK();
}

void vec_copy(const double *x, int64_t L, double *y) {
    SVB_CUDA(cudaMemcpyAsync(y, x, (size_t)L * sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream));
}

// ---- counter-based normals (Philox4x32-10 + Box-Muller) ----------------------------------------------
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
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

__global__ void fill_normal_kernel(double *x, int64_t L, uint64_t seed, uint64_t offset) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t ctr = (uint64_t)i + offset;
        uint32_t r[4];
        philox4x32((uint32_t)ctr, (uint32_t)(ctr >> 32), 0x5eed0001u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        const double u1 = ((double)(((uint64_t)r[0] << 21) ^ (r[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);  // (0,1)
        const double u2 = ((double)(((uint64_t)r[2] << 21) ^ (r[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
        x[i] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
}

void vec_fill_normal(double *x, int64_t L, uint64_t seed, uint64_t offset) {
    const int nrb = row_blocks(L, 256, 8);
    KTimer kt(SVB_K_VECTOR, 8.0 * L);
    fill_normal_kernel<<<(unsigned)nrb, 256, 0, ctx().stream>>>(x, L, seed, offset);
    SVB_LAUNCH_CHECK();
}

}  // namespace svb

extern "C" int svb_synth_normal(int64_t n, uint64_t seed, double *host_out) {
    SVB_API_BEGIN
    svb::require_init();
    SVB_CHECK(host_out && n >= 0, SVB_EARG, "svb_synth_normal: bad argument");
    svb::DevBuf<double> d((size_t)std::max<int64_t>(n, 1));
    svb::vec_fill_normal(d.p, n, seed, 0);
    SVB_CUDA(cudaMemcpyAsync(host_out, d.p, (size_t)n * 8, cudaMemcpyDeviceToHost, svb::ctx().stream));
    SVB_CUDA(cudaStreamSynchronize(svb::ctx().stream));
    SVB_API_END
}

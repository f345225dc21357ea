// spmm.cu — the matrix forms of the centred operator (scaling.jl:259-272): Y = alpha*S*X + beta*Y and
// Y = alpha*S'*X + beta*Y for k right-hand sides. Same two streaming layouts, same persistent equal-nnz
// partition and dynamic segment grabbing as the SpMV kernels (operator.cu); KC = 4 right-hand sides are
// carried per pass, so every nonzero is read once per 4 columns (bytes per column / 4). The gathered operand
// (X rows: n x KC, or the W tile: R x KC) is staged row-major in shared memory so one nonzero reads KC
// consecutive doubles. The adjoint form SUBTRACTS the rank-1 term alpha*mu*sum(X) — scaling.jl:271 adds it,
// an untested sign slip upstream; the vector form :256 and the maths say subtract.
#include "svb_internal.h"

#include <algorithm>

namespace svb {
namespace {

constexpr int KC = 4;

__device__ __forceinline__ double block_sum_mm(double v, double *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

__device__ __forceinline__ int64_t lower_bound_mm(const int64_t *__restrict__ ptr, int64_t len, int64_t target) {
    int64_t lo = 0, hi = len;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(ptr + mid) < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void cta_range_mm(const int64_t *__restrict__ ptr, int64_t nseg, int64_t nnz, int64_t &s0, int64_t &s1) {
    const int64_t G = gridDim.x, b = blockIdx.x;
    const int64_t lo = (int64_t)(((__int128)nnz * b) / G), hi = (int64_t)(((__int128)nnz * (b + 1)) / G);
    s0 = (b == 0) ? 0 : lower_bound_mm(ptr, nseg, lo);
    s1 = (b == G - 1) ? nseg : lower_bound_mm(ptr, nseg, hi);
}

// acc[c] += sum_k val[k] * xs[idx[k]*KC + c] over one segment, reduced over the LPS lanes of the sub-warp
template <typename V, typename IdxT, int LPS>
__device__ __forceinline__ void seg_dot_mm(const V *__restrict__ val, const IdxT *__restrict__ idx, int64_t beg, int64_t end,
                                           const double *__restrict__ xs, int sub_lane, unsigned submask, double acc[KC]) {
    double a[KC], b[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) { a[c] = 0.0; b[c] = 0.0; }
    int64_t k = beg + sub_lane;
    for (; k + LPS < end; k += 2 * LPS) {
        const double v0 = (double)__ldg(val + k), v1 = (double)__ldg(val + k + LPS);
        const double *x0 = xs + (size_t)__ldg(idx + k) * KC, *x1 = xs + (size_t)__ldg(idx + k + LPS) * KC;
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            a[c] = fma(v0, x0[c], a[c]);
            b[c] = fma(v1, x1[c], b[c]);
        }
    }
    if (k < end) {
        const double v0 = (double)__ldg(val + k);
        const double *x0 = xs + (size_t)__ldg(idx + k) * KC;
#pragma unroll
        for (int c = 0; c < KC; ++c) a[c] = fma(v0, x0[c], a[c]);
    }
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        double s = a[c] + b[c];
#pragma unroll
        for (int o = LPS >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(submask, s, o);
        acc[c] = s;
    }
}

template <typename V, typename IdxT, int LPS, bool XSMEM>
__global__ void __launch_bounds__(1024) spmm_fwd_kernel(const int64_t *__restrict__ rowptr, const IdxT *__restrict__ fidx,
                                                        const V *__restrict__ fval, int64_t m, int64_t n, int64_t nnz,
                                                        const double *__restrict__ X, int64_t ldx, int kc,
                                                        const double *__restrict__ mu, double alpha, double beta,
                                                        double *__restrict__ Y, int64_t ldy, const double *__restrict__ xrow_global) {
    extern __shared__ double smem[];
    __shared__ unsigned long long next_row;
    double *red = smem;      // 32
    double *xs = smem + 32;  // n*KC when XSMEM
    double part[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) part[c] = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            const double xv = (c < kc) ? X[j + (int64_t)c * ldx] : 0.0;
            if (XSMEM) xs[j * KC + c] = xv;
            if (mu) part[c] = fma(mu[j], xv, part[c]);
        }
    }
    int64_t r0, r1;
    cta_range_mm(rowptr, m, nnz, r0, r1);
    const int nsub = blockDim.x / LPS;
    if (threadIdx.x == 0) next_row = (unsigned long long)(r0 + nsub);
    double mudot[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) mudot[c] = block_sum_mm(part[c], red);
    const double *xg = XSMEM ? xs : xrow_global;  // row-major n x KC copy in global memory when x does not fit
    const int lane = threadIdx.x & 31;
    const int sub_lane = lane & (LPS - 1);
    const unsigned submask = (LPS == 32) ? 0xffffffffu : (((1u << LPS) - 1u) << (lane & ~(LPS - 1)));
    int64_t row = r0 + threadIdx.x / LPS;
    while (row < r1) {
        double acc[KC];
        seg_dot_mm<V, IdxT, LPS>(fval, fidx, __ldg(rowptr + row), __ldg(rowptr + row + 1), xg, sub_lane, submask, acc);
        unsigned long long nxt = 0;
        if (sub_lane == 0) {
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                if (c < kc) {
                    double r = alpha * (acc[c] - mudot[c]);
                    if (beta != 0.0) r = fma(beta, Y[row + (int64_t)c * ldy], r);
                    Y[row + (int64_t)c * ldy] = r;
                }
            }
            nxt = atomicAdd(&next_row, 1ull);
        }
        row = (int64_t)__shfl_sync(submask, nxt, lane & ~(LPS - 1));
    }
}

template <typename V, int LPS>
__global__ void __launch_bounds__(1024) spmm_adj_kernel(const int64_t *__restrict__ gptr, const uint16_t *__restrict__ rloc,
                                                        const V *__restrict__ aval, int64_t m, int64_t n, int log2R,
                                                        int64_t ntiles, int64_t nnz, const double *__restrict__ W, int64_t ldw,
                                                        int kc, double *__restrict__ partial /* [ntiles][n+1][KC] */) {
    extern __shared__ double smem[];
    __shared__ unsigned long long next_seg;
    double *red = smem;
    double *ws = smem + 32;  // R*KC
    const int64_t R = (int64_t)1 << log2R;
    int64_t s0, s1;
    cta_range_mm(gptr, ntiles * n, nnz, s0, s1);
    const int nsub = blockDim.x / LPS;
    const int lane = threadIdx.x & 31;
    const int sub_lane = lane & (LPS - 1);
    const unsigned submask = (LPS == 32) ? 0xffffffffu : (((1u << LPS) - 1u) << (lane & ~(LPS - 1)));
    for (int64_t t = s0 / n; t < ntiles && t * n < s1; ++t) {
        const int64_t a = max(s0, t * n), b = min(s1, (t + 1) * n);
        const int64_t row0 = t << log2R;
        __syncthreads();
        double part[KC];
#pragma unroll
        for (int c = 0; c < KC; ++c) part[c] = 0.0;
        for (int64_t r = threadIdx.x; r < R; r += blockDim.x) {
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                const double wv = (row0 + r < m && c < kc) ? W[row0 + r + (int64_t)c * ldw] : 0.0;
                ws[r * KC + c] = wv;
                part[c] += wv;
            }
        }
        if (threadIdx.x == 0) next_seg = (unsigned long long)(a + nsub);
        double wsum[KC];
#pragma unroll
        for (int c = 0; c < KC; ++c) wsum[c] = block_sum_mm(part[c], red);
        double *pt = partial + (size_t)t * (n + 1) * KC;
        if (a == t * n && threadIdx.x == 0) {
#pragma unroll
            for (int c = 0; c < KC; ++c) pt[(size_t)n * KC + c] = wsum[c];
        }
        int64_t s = a + threadIdx.x / LPS;
        while (s < b) {
            double acc[KC];
            seg_dot_mm<V, uint16_t, LPS>(aval, rloc, __ldg(gptr + s), __ldg(gptr + s + 1), ws, sub_lane, submask, acc);
            unsigned long long nxt = 0;
            if (sub_lane == 0) {
#pragma unroll
                for (int c = 0; c < KC; ++c) pt[(size_t)(s - t * n) * KC + c] = acc[c];
                nxt = atomicAdd(&next_seg, 1ull);
            }
            s = (int64_t)__shfl_sync(submask, nxt, lane & ~(LPS - 1));
        }
    }
}

// tmp[g + c*n] = sum_t partial[t][g][c] - (sum_t partial[t][n][c]) * mu[g]   (fixed order); when `final`:
// Y[g + c*ldy] = alpha*tmp + beta*Y
__global__ void __launch_bounds__(256) spmm_adj_reduce_kernel(const double *__restrict__ partial, int64_t ntiles, int64_t n, int kc,
                                                              const double *__restrict__ mu, double *__restrict__ tmp, int final,
                                                              double alpha, double beta, double *__restrict__ Y, int64_t ldy) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    for (int c = 0; c < kc; ++c) {
        double s = 0.0, ws = 0.0;
        for (int64_t t = 0; t < ntiles; ++t) {
            const double *pt = partial + (size_t)t * (n + 1) * KC;
            s += pt[(size_t)g * KC + c];
            ws += pt[(size_t)n * KC + c];
        }
        const double v = mu ? fma(-ws, mu[g], s) : s;
        if (final) {
            double r = alpha * v;
            if (beta != 0.0) r = fma(beta, Y[g + (int64_t)c * ldy], r);
            Y[g + (int64_t)c * ldy] = r;
        } else {
            tmp[g + (int64_t)c * n] = v;
        }
    }
}

__global__ void spmm_combine_kernel(int64_t n, int kc, double alpha, const double *__restrict__ tmp, double beta,
                                    double *__restrict__ Y, int64_t ldy) {
    const int64_t total = n * kc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = i / n, g = i - c * n;
        double r = alpha * tmp[i];
        if (beta != 0.0) r = fma(beta, Y[g + c * ldy], r);
        Y[g + c * ldy] = r;
    }
}

__global__ void pack_rows_kernel(const double *__restrict__ X, int64_t ldx, int64_t n, int kc, double *__restrict__ out) {
    const int64_t total = n * KC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = i / KC;
        const int c = (int)(i - j * KC);
        out[i] = (c < kc) ? X[j + (int64_t)c * ldx] : 0.0;
    }
}

template <typename K>
int resident_grid_mm(K kernel, int threads, size_t smem) {
    int per_sm = 0;
    SVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    return std::max(1, per_sm) * ctx().sm_count;
}

template <typename V, typename IdxT, int LPS>
void launch_spmm_fwd(svb_operator_s *op, double alpha, const double *dX, int64_t ldx, int kc, double beta, double *dY, int64_t ldy) {
    Context &C = ctx();
    const size_t xs_bytes = (size_t)op->n * KC * sizeof(double);
    const bool xsmem = xs_bytes + 2048 <= C.smem_optin;
    const size_t smem = 32 * sizeof(double) + (xsmem ? xs_bytes : 0);
    const int threads = smem > 64 * 1024 ? 1024 : 256;
    if (xsmem) {
        auto k = spmm_fwd_kernel<V, IdxT, LPS, true>;
        if (smem > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = (int)std::min<int64_t>(resident_grid_mm(k, threads, smem), std::max<int64_t>(1, op->m / 8));
        k<<<grid, threads, smem, C.stream>>>(op->rowptr, (const IdxT *)op->fidx, (const V *)op->fval, op->m, op->n, op->nnz, dX, ldx, kc,
                                             op->mu, alpha, beta, dY, ldy, nullptr);
    } else {
        DevBuf<double> xr((size_t)op->n * KC);
        pack_rows_kernel<<<(unsigned)std::min<int64_t>((op->n * KC + 255) / 256, 148 * 8), 256, 0, C.stream>>>(dX, ldx, op->n, kc, xr.p);
        auto k = spmm_fwd_kernel<V, IdxT, LPS, false>;
        const int grid = (int)std::min<int64_t>(resident_grid_mm(k, threads, smem), std::max<int64_t>(1, op->m / 8));
        k<<<grid, threads, smem, C.stream>>>(op->rowptr, (const IdxT *)op->fidx, (const V *)op->fval, op->m, op->n, op->nnz, dX, ldx, kc,
                                             op->mu, alpha, beta, dY, ldy, xr.p);
        SVB_CUDA(cudaStreamSynchronize(C.stream));  // xr is freed on return
    }
    SVB_LAUNCH_CHECK();
}

template <typename V, int LPS>
void launch_spmm_adj(svb_operator_s *op, const double *dW, int64_t ldw, int kc, double *partial) {
    Context &C = ctx();
    const size_t smem = (32 + (size_t)op->R * KC) * sizeof(double);
    SVB_CHECK(smem + 1024 <= C.smem_optin, SVB_EDIM, "spmm adjoint: tile does not fit in shared memory");
    auto k = spmm_adj_kernel<V, LPS>;
    if (smem > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = smem > 64 * 1024 ? 1024 : 256;
    const int64_t nseg = op->ntiles * op->n;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(resident_grid_mm(k, threads, smem), nseg / 8 + 1));
    k<<<grid, threads, smem, C.stream>>>(op->gptr, op->rloc, (const V *)op->aval, op->m, op->n, (int)op->log2R, op->ntiles, op->nnz, dW,
                                         ldw, kc, partial);
    SVB_LAUNCH_CHECK();
}


// ---------------------------------------------------------------------------------------------
// Gram matrix C'C (scaling.jl:274-296, with the CSC x CSC product of mul.jl:82-114 underneath) — the n x n input of
// `tssvd` (embedding.jl:30-44). One pass per block of KC genes J = [j0, j0+kc) over the adjoint layout ALONE: a CTA
// densifies the KC columns of its cell tile into shared memory straight from their (tile, gene) segments (R x KC doubles,
// no m x KC intermediate in HBM, no forward product), then every segment (tile, g) with g >= j0 is dotted against the KC
// dense columns — the same gather loop as the adjoint SpMM. Only the lower triangle is computed (symmetry halves the
// stream); per-tile partials are reduced in tile order (deterministic), written to G[g, j] and mirrored to G[j, g].
// ---------------------------------------------------------------------------------------------
template <typename V, int LPS>
__global__ void __launch_bounds__(1024) gram_adj_kernel(const int64_t *__restrict__ gptr, const uint16_t *__restrict__ rloc,
                                                        const V *__restrict__ aval, int64_t n, int log2R, int64_t ntiles, int64_t nnz,
                                                        int64_t j0, int kc, double *__restrict__ partial /* [ntiles][n][KC] */) {
    extern __shared__ double smem[];
    __shared__ unsigned long long next_seg;
    double *ws = smem;  // R*KC: the dense columns J of this cell tile, row-major
    const int64_t R = (int64_t)1 << log2R;
    int64_t s0, s1;
    cta_range_mm(gptr, ntiles * n, nnz, s0, s1);
    const int nsub = blockDim.x / LPS;
    const int lane = threadIdx.x & 31;
    const int sub_lane = lane & (LPS - 1);
    const unsigned submask = (LPS == 32) ? 0xffffffffu : (((1u << LPS) - 1u) << (lane & ~(LPS - 1)));
    for (int64_t t = s0 / n; t < ntiles && t * n < s1; ++t) {
        const int64_t a = max(max(s0, t * n), t * n + j0), b = min(s1, (t + 1) * n);
        if (a >= b) continue;  // uniform over the CTA
        __syncthreads();       // every sub-warp has left the previous tile (ws, next_seg)
        for (int64_t r = threadIdx.x; r < R * KC; r += blockDim.x) ws[r] = 0.0;
        __syncthreads();
        for (int c = 0; c < kc; ++c) {
            const int64_t seg = t * n + j0 + c;
            const int64_t k1 = __ldg(gptr + seg + 1);
            for (int64_t k = __ldg(gptr + seg) + threadIdx.x; k < k1; k += blockDim.x)
                ws[(size_t)__ldg(rloc + k) * KC + c] = (double)__ldg(aval + k);
        }
        if (threadIdx.x == 0) next_seg = (unsigned long long)(a + nsub);
        __syncthreads();
        double *pt = partial + (size_t)t * n * KC;
        int64_t s = a + threadIdx.x / LPS;
        while (s < b) {
            double acc[KC];
            seg_dot_mm<V, uint16_t, LPS>(aval, rloc, __ldg(gptr + s), __ldg(gptr + s + 1), ws, sub_lane, submask, acc);
            unsigned long long nxt = 0;
            if (sub_lane == 0) {
#pragma unroll
                for (int c = 0; c < KC; ++c) pt[(size_t)(s - t * n) * KC + c] = acc[c];
                nxt = atomicAdd(&next_seg, 1ull);
            }
            s = (int64_t)__shfl_sync(submask, nxt, lane & ~(LPS - 1));
        }
    }
}

// G[g, j0+c] = G[j0+c, g] = sum_t partial[t][g][c] for g >= j0+c (tile order fixed). Tiles whose segments of gene g hold no
// nonzero still wrote a zero partial (every segment >= j0 of every tile is visited by exactly one sub-warp).
__global__ void __launch_bounds__(256) gram_reduce_kernel(const double *__restrict__ partial, int64_t ntiles, int64_t n, int64_t j0,
                                                          int kc, double *__restrict__ G, int64_t ldg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t g = j0 + i / KC;
    const int c = (int)(i % KC);
    if (g >= n || c >= kc || g < j0 + c) return;
    double s = 0.0;
    for (int64_t t = 0; t < ntiles; ++t) s += partial[((size_t)t * n + g) * KC + c];
    G[g + (j0 + c) * ldg] = s;
    G[(j0 + c) + g * ldg] = s;
}

// G += -mu q' - q mu' + M mu mu'   (scaling.jl:281-294; q = column sums of A, M = cells of the whole matrix). Products are
// rounded separately so that the update — and therefore G — stays exactly symmetric.
__global__ void __launch_bounds__(256) gram_centre_kernel(double *__restrict__ G, int64_t n, int64_t ldg, const double *__restrict__ mu,
                                                          const double *__restrict__ q, double M) {
    const int64_t total = n * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / n, a = i - b * n;
        const double t = __dadd_rn(__dmul_rn(mu[a], q[b]), __dmul_rn(q[a], mu[b]));
        const double u = __dmul_rn(M, __dmul_rn(mu[a], mu[b]));
        G[a + b * ldg] = __dadd_rn(G[a + b * ldg], __dadd_rn(u, -t));
    }
}

__global__ void fill_kernel(double *__restrict__ x, int64_t L, double v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}

template <typename V, int LPS>
void launch_gram_adj(svb_operator_s *op, int64_t j0, int kc, double *partial) {
    Context &C = ctx();
    const size_t smem = (size_t)op->R * KC * sizeof(double);
    SVB_CHECK(smem + 1024 <= C.smem_optin, SVB_EDIM, "gram: tile does not fit in shared memory");
    auto k = gram_adj_kernel<V, LPS>;
    if (smem > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = smem > 64 * 1024 ? 1024 : 256;
    const int64_t nseg = op->ntiles * op->n;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(resident_grid_mm(k, threads, smem), nseg / 8 + 1));
    k<<<grid, threads, smem, C.stream>>>(op->gptr, op->rloc, (const V *)op->aval, op->n, (int)op->log2R, op->ntiles, op->nnz, j0, kc, partial);
    count_launch();
    SVB_LAUNCH_CHECK();
}

__global__ void __launch_bounds__(256) symmetrise_kernel(double *__restrict__ G, int64_t n, int64_t ldg) {
    const int64_t total = n * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / n, a = i - b * n;
        if (a < b) {
            const double v = 0.5 * (G[a + b * ldg] + G[b + a * ldg]);
            G[a + b * ldg] = v;
            G[b + a * ldg] = v;
        }
    }
}

unsigned grid1d_mm(int64_t n, int threads = 256) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, 148 * 8));
}

}  // namespace

// Y (outL x k, ldy) = alpha * op(S) * X (inL x k, ldx) + beta * Y ; device pointers, sparse operators only
void op_apply_mm(svb_operator_s *op, bool trans, double alpha, const double *dX, int64_t ldx, double beta, double *dY, int64_t ldy,
                 int64_t k) {
    Context &C = ctx();
    if (op->fact) {
        // count-level operator: its matrix forms are column-by-column vector products (same 2 B/nnz stream per column)
        for (int64_t c = 0; c < k; ++c) op_apply(op, trans, alpha, dX + c * ldx, beta, dY + c * ldy);
        return;
    }
    const double favg = op->m > 0 ? (double)op->nnz / (double)op->m : 0.0;
    const double aavg = (op->n > 0 && op->ntiles > 0) ? (double)op->nnz / ((double)op->n * (double)op->ntiles) : 0.0;
    DevBuf<double> partial;
    if (trans) partial.alloc((size_t)op->ntiles * (op->n + 1) * KC);
    DevBuf<double> tmp;
    const bool multi = C.nranks > 1;
    if (trans && multi) tmp.alloc((size_t)op->n * KC);
    for (int64_t c0 = 0; c0 < k; c0 += KC) {
        const int kc = (int)std::min<int64_t>(KC, k - c0);
        const double *Xc = dX + c0 * ldx;
        double *Yc = dY + c0 * ldy;
        if (!trans) {
            KTimer kt(SVB_K_SPMV_FWD, (double)op->nnz * (op->vbytes + op->ibytes) + 8.0 * (op->m + 1) + 8.0 * kc * (op->n + op->m));
            const bool wide = favg >= 64;
#define SVB_FWD_MM(V, I) (wide ? launch_spmm_fwd<V, I, 32>(op, alpha, Xc, ldx, kc, beta, Yc, ldy) : launch_spmm_fwd<V, I, 8>(op, alpha, Xc, ldx, kc, beta, Yc, ldy))
            if (op->vbytes == 8) {
                if (op->ibytes == 2) SVB_FWD_MM(double, uint16_t); else SVB_FWD_MM(double, int32_t);
            } else {
                if (op->ibytes == 2) SVB_FWD_MM(float, uint16_t); else SVB_FWD_MM(float, int32_t);
            }
#undef SVB_FWD_MM
        } else {
            {
                KTimer kt(SVB_K_SPMV_ADJ, (double)op->nnz * (op->vbytes + 2) + 8.0 * ((double)op->ntiles * op->n + 1) + 8.0 * kc * (op->n + op->m), 2);
                const bool wide = aavg >= 64;
                if (op->vbytes == 8) {
                    if (wide) launch_spmm_adj<double, 32>(op, Xc, ldx, kc, partial.p); else launch_spmm_adj<double, 8>(op, Xc, ldx, kc, partial.p);
                } else {
                    if (wide) launch_spmm_adj<float, 32>(op, Xc, ldx, kc, partial.p); else launch_spmm_adj<float, 8>(op, Xc, ldx, kc, partial.p);
                }
                spmm_adj_reduce_kernel<<<(unsigned)((op->n + 255) / 256), 256, 0, C.stream>>>(partial.p, op->ntiles, op->n, kc, op->mu, tmp.p,
                                                                                         multi ? 0 : 1, alpha, beta, Yc, ldy);
                SVB_LAUNCH_CHECK();
            }
            if (multi) {
                comm_allreduce_dev(tmp.p, op->n * kc);
                spmm_combine_kernel<<<(unsigned)std::min<int64_t>((op->n * kc + 255) / 256, 148 * 8), 256, 0, C.stream>>>(op->n, kc, alpha, tmp.p,
                                                                                                                beta, Yc, ldy);
                count_launch();
                SVB_LAUNCH_CHECK();
            }
        }
    }
    SVB_CUDA(cudaStreamSynchronize(C.stream));  // partial / tmp are freed on return
}

// Y (outL x k) = op(S) * X (inL x k) for ANY operator kind (dense operators have no SpMM form: column by column)
void op_apply_cols(svb_operator_s *op, bool trans, const double *dX, int64_t ldx, double *dY, int64_t ldy, int64_t k) {
    if (op->dense) {
        for (int64_t c = 0; c < k; ++c) op_apply(op, trans, 1.0, dX + c * ldx, 0.0, dY + c * ldy);
        return;
    }
    op_apply_mm(op, trans, 1.0, dX, ldx, 0.0, dY, ldy, k);
}

// G (n x n, column-major, leading dimension n, device) = S'S of the centred operator — `C'C`, scaling.jl:274-296.
// Explicit sparse operators: A'A by the fused tile kernels above (summed over the ranks when cells are sharded), then the
// rank-1 terms of scaling.jl:281-294. Dense and count-level operators: column j = S'(S e_j) through the vector products.
void op_gram(svb_operator_s *op, double *G) {
    Context &C = ctx();
    cudaStream_t st = C.stream;
    const int64_t n = op->n, m = op->m;
    if (op->dense || op->fact) {
        DevBuf<double> e((size_t)n), y((size_t)m);
        SVB_CUDA(cudaMemsetAsync(e.p, 0, (size_t)n * 8, st));
        for (int64_t j = 0; j < n; ++j) {
            fill_kernel<<<1, 32, 0, st>>>(e.p + j, 1, 1.0);
            op_apply(op, false, 1.0, e.p, 0.0, y.p);
            op_apply(op, true, 1.0, y.p, 0.0, G + j * n);
            fill_kernel<<<1, 32, 0, st>>>(e.p + j, 1, 0.0);
            count_launch(2);
        }
        symmetrise_kernel<<<grid1d_mm(n * n), 256, 0, st>>>(G, n, n);
        count_launch();
        SVB_LAUNCH_CHECK();
        SVB_CUDA(cudaStreamSynchronize(st));
        return;
    }
    {
        DevBuf<double> partial((size_t)op->ntiles * n * KC);
        const double aavg = (n > 0 && op->ntiles > 0) ? (double)op->nnz / ((double)n * (double)op->ntiles) : 0.0;
        const bool wide = aavg >= 64;
        for (int64_t j0 = 0; j0 < n; j0 += KC) {
            const int kc = (int)std::min<int64_t>(KC, n - j0);
            if (op->vbytes == 8) {
                if (wide) launch_gram_adj<double, 32>(op, j0, kc, partial.p); else launch_gram_adj<double, 8>(op, j0, kc, partial.p);
            } else {
                if (wide) launch_gram_adj<float, 32>(op, j0, kc, partial.p); else launch_gram_adj<float, 8>(op, j0, kc, partial.p);
            }
            gram_reduce_kernel<<<(unsigned)(((n - j0) * KC + 255) / 256), 256, 0, st>>>(partial.p, op->ntiles, n, j0, kc, G, n);
            count_launch();
            SVB_LAUNCH_CHECK();
        }
        SVB_CUDA(cudaStreamSynchronize(st));  // partial is freed here
    }
    if (C.nranks > 1) comm_allreduce_dev(G, n * n);
    if (op->mu) {
        // q = A'1 (the operator with its centre switched off), M = cells of the whole matrix
        DevBuf<double> ones((size_t)m), q((size_t)n), md(1);
        fill_kernel<<<grid1d_mm(m), 256, 0, st>>>(ones.p, m, 1.0);
        count_launch();
        double *mu = op->mu;
        op->mu = nullptr;
        try {
            op_apply(op, true, 1.0, ones.p, 0.0, q.p);
        } catch (...) {
            op->mu = mu;
            throw;
        }
        op->mu = mu;
        double M = (double)m;
        if (C.nranks > 1) {
            SVB_CUDA(cudaMemcpyAsync(md.p, &M, 8, cudaMemcpyHostToDevice, st));
            comm_allreduce_dev(md.p, 1);
            SVB_CUDA(cudaMemcpyAsync(&M, md.p, 8, cudaMemcpyDeviceToHost, st));
            SVB_CUDA(cudaStreamSynchronize(st));
        }
        gram_centre_kernel<<<grid1d_mm(n * n), 256, 0, st>>>(G, n, n, op->mu, q.p, M);
        count_launch();
        SVB_LAUNCH_CHECK();
        SVB_CUDA(cudaStreamSynchronize(st));  // ones / q are freed here
    }
}

}  // namespace svb

extern "C" {

int svb_gram(svb_operator_t op, double *G) {
    SVB_API_BEGIN
    svb::require_init();
    SVB_CHECK(op && G, SVB_EARG, "svb_gram: null argument");
    svb::DevBuf<double> dG((size_t)op->n * op->n);
    svb::op_gram(op, dG.p);
    SVB_CUDA(cudaMemcpyAsync(G, dG.p, (size_t)op->n * op->n * 8, cudaMemcpyDeviceToHost, svb::ctx().stream));
    SVB_CUDA(cudaStreamSynchronize(svb::ctx().stream));
    SVB_API_END
}

}  // extern "C"

// operator.cu — the implicit centred operator S = A - 1*mu' (scaling.jl:219-232) on the device and
// its two products:
//   forward  y = alpha*S*x  + beta*y  (scaling.jl:245-250 : mul!(C, S, v, a, b)  + stdlib CSC*vec)
//   adjoint  y = alpha*S'*x + beta*y  (scaling.jl:252-257 : mul!(C, S', v, a, b) + stdlib CSC'*vec)
// Both stream the nonzeros exactly once from HBM in a layout made for that product:
//   forward : CSR by cell, gene index u16 (n <= 65535) or i32; x (n doubles) lives in shared memory,
//             a sub-warp of LPS lanes owns one cell, warp-shuffle reduction, mu.x fused.
//   adjoint : cells tiled by R; inside a tile nonzeros are gene-major with a u16 local cell index;
//             the w tile (R doubles) lives in shared memory, a sub-warp owns one (tile, gene)
//             segment, per-tile partials are reduced in a fixed order (deterministic), the rank-1
//             term -(sum w)*mu and an optional axpy are fused in the reduce epilogue.
#include "svb_internal.h"
#include "layout.cuh"
#include "p2p.cuh"

#include <algorithm>
#include <cstdlib>

using namespace svb;

svb_operator_s::~svb_operator_s() {
    void *ptrs[] = {mu, rowptr, fidx, fval, gptr, rloc, aval, partial, dA, xdev, ydev, tmp, scal, fwd_ranges, adj_ranges};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    delete fact;
}

namespace svb {

static int env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return s ? atoi(s) : dflt;
}

// ---------------------------------------------------------------------------------------------
// shared device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double *red /* >= 32 doubles */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += red[i];  // fixed order: same value in every thread
    return t;
}

template <typename V, typename IdxT, int LPS>
__device__ __forceinline__ double seg_dot(const V *__restrict__ val, const IdxT *__restrict__ idx, int64_t beg,
                                          int64_t end, const double *__restrict__ xs, int sub_lane, unsigned submask) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int64_t k = beg + sub_lane;
    for (; k + 3 * LPS < end; k += 4 * LPS) {
        const V v0 = __ldg(val + k), v1 = __ldg(val + k + LPS), v2 = __ldg(val + k + 2 * LPS), v3 = __ldg(val + k + 3 * LPS);
        const IdxT i0 = __ldg(idx + k), i1 = __ldg(idx + k + LPS), i2 = __ldg(idx + k + 2 * LPS), i3 = __ldg(idx + k + 3 * LPS);
        a0 = fma((double)v0, xs[i0], a0);
        a1 = fma((double)v1, xs[i1], a1);
        a2 = fma((double)v2, xs[i2], a2);
        a3 = fma((double)v3, xs[i3], a3);
    }
    for (; k < end; k += LPS) a0 = fma((double)__ldg(val + k), xs[__ldg(idx + k)], a0);
    double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = LPS >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(submask, acc, o);
    return acc;
}

// first index s in [0, len] with ptr[s] >= target (ptr ascending)
__device__ __forceinline__ int64_t lower_bound_i64(const int64_t *__restrict__ ptr, int64_t len, int64_t target) {
    int64_t lo = 0, hi = len;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(ptr + mid) < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// CTA b of G owns the segments whose first nonzero lies in its 1/G share of the nonzero stream: equal
// bytes per CTA whatever the segment lengths are. Every segment id belongs to exactly one CTA. The G+1
// boundaries are computed once per operator (two dependent 20-step binary searches per CTA cost ~25 us per
// launch, 13 % of a product once the cells are sharded over 8 GPUs).
__global__ void cta_ranges_kernel(const int64_t *__restrict__ ptr, int64_t nseg, int64_t nnz, int G, int64_t *__restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > G) return;
    if (b == 0) out[0] = 0;
    else if (b == G) out[G] = nseg;
    else out[b] = lower_bound_i64(ptr, nseg, (int64_t)(((__int128)nnz * b) / G));
}

// ---------------------------------------------------------------------------------------------
// forward: y_i = alpha*(sum_j a_ij x_j - mu.x) + beta*y_i + csign*(*coef)*cvec_i
// persistent grid (one resident wave); sub-warps of LPS lanes grab cells dynamically inside the CTA's range
// ---------------------------------------------------------------------------------------------
template <typename V, typename IdxT, int LPS, bool XSMEM, int BLOCK>
__global__ void __launch_bounds__(BLOCK, (BLOCK == 256 ? 6 : 3)) spmv_fwd_kernel(const int64_t *__restrict__ rowptr, const IdxT *__restrict__ fidx,
                                                          const V *__restrict__ fval, int64_t m, int64_t n, int64_t nnz,
                                                          const double *__restrict__ x, const double *__restrict__ mu,
                                                          double alpha, double beta, double *__restrict__ y,
                                                          const double *__restrict__ coef, double csign,
                                                          const double *__restrict__ cvec, const int64_t *__restrict__ ranges) {
    extern __shared__ double smem[];
    __shared__ unsigned long long next_row;
    double *red = smem;       // 32 doubles
    double *xs = smem + 32;   // n doubles when XSMEM
    double part = 0.0;
    if (XSMEM) {
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
            const double xv = x[j];
            xs[j] = xv;
            if (mu) part = fma(mu[j], xv, part);
        }
    } else if (mu) {
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) part = fma(mu[j], x[j], part);
    }
    const int64_t r0 = __ldg(ranges + blockIdx.x), r1 = __ldg(ranges + blockIdx.x + 1);
    constexpr int NSUB = BLOCK / LPS;
    if (threadIdx.x == 0) next_row = (unsigned long long)(r0 + NSUB);
    const double mudot = mu ? block_sum(part, red) : 0.0;  // contains the __syncthreads that publish xs / next_row
    if (!mu) __syncthreads();
    const double *xg = XSMEM ? xs : x;
    const double c = (coef != nullptr) ? csign * (*coef) : 0.0;

    const int lane = threadIdx.x & 31;
    const int sub_lane = lane & (LPS - 1);
    const unsigned submask = (LPS == 32) ? 0xffffffffu : (((1u << LPS) - 1u) << (lane & ~(LPS - 1)));
    int64_t row = r0 + threadIdx.x / LPS;
    while (row < r1) {
        const double acc = seg_dot<V, IdxT, LPS>(fval, fidx, __ldg(rowptr + row), __ldg(rowptr + row + 1), xg, sub_lane, submask);
        unsigned long long nxt = 0;
        if (sub_lane == 0) {
            double r = alpha * (acc - mudot);
            if (beta != 0.0) r = fma(beta, y[row], r);
            if (coef != nullptr) r = fma(c, cvec[row], r);
            y[row] = r;
            nxt = atomicAdd(&next_row, 1ull);
        }
        row = (int64_t)__shfl_sync(submask, nxt, lane & ~(LPS - 1));
    }
}

// ---------------------------------------------------------------------------------------------
// adjoint, stage 1: partial[t][g] = sum_{cells i in tile t} a_ig * w_i ; partial[t][n] = sum_i w_i
// persistent grid; a CTA walks its equal-nnz range of (tile, gene) segments, reloading the w tile
// (R doubles, shared memory) when the range crosses a tile boundary.
// ---------------------------------------------------------------------------------------------
template <typename V, int LPS>
__global__ void __launch_bounds__(256, 5) spmv_adj_kernel(const int64_t *__restrict__ gptr, const uint16_t *__restrict__ rloc,
                                                          const V *__restrict__ aval, int64_t m, int64_t n, int log2R,
                                                          int64_t ntiles, int64_t nnz, const double *__restrict__ w,
                                                          double *__restrict__ partial, const int64_t *__restrict__ ranges) {
    extern __shared__ double smem[];
    __shared__ unsigned long long next_seg;
    double *red = smem;      // 32
    double *ws = smem + 32;  // R
    const int64_t R = (int64_t)1 << log2R;
    const int64_t s0 = __ldg(ranges + blockIdx.x), s1 = __ldg(ranges + blockIdx.x + 1);
    constexpr int NSUB = 256 / LPS;
    const int lane = threadIdx.x & 31;
    const int sub_lane = lane & (LPS - 1);
    const unsigned submask = (LPS == 32) ? 0xffffffffu : (((1u << LPS) - 1u) << (lane & ~(LPS - 1)));
    for (int64_t t = s0 / n; t < ntiles && t * n < s1; ++t) {
        const int64_t a = max(s0, t * n), b = min(s1, (t + 1) * n);
        const int64_t row0 = t << log2R;
        __syncthreads();  // everybody is done with the previous tile
        double part = 0.0;
        for (int64_t r = threadIdx.x; r < R; r += blockDim.x) {
            const double wv = (row0 + r < m) ? w[row0 + r] : 0.0;
            ws[r] = wv;
            part += wv;
        }
        if (threadIdx.x == 0) next_seg = (unsigned long long)(a + NSUB);
        const double wsum = block_sum(part, red);  // also publishes ws / next_seg
        if (a == t * n && threadIdx.x == 0) partial[t * (n + 1) + n] = wsum;
        int64_t s = a + threadIdx.x / LPS;
        while (s < b) {
            const double acc = seg_dot<V, uint16_t, LPS>(aval, rloc, __ldg(gptr + s), __ldg(gptr + s + 1), ws, sub_lane, submask);
            unsigned long long nxt = 0;
            if (sub_lane == 0) {
                partial[t * (n + 1) + (s - t * n)] = acc;
                nxt = atomicAdd(&next_seg, 1ull);
            }
            s = (int64_t)__shfl_sync(submask, nxt, lane & ~(LPS - 1));
        }
    }
}

// adjoint, stage 2: tmp[g] = sum_t partial[t][g] - (sum_t partial[t][n]) * mu[g]   (fixed order)
// and, when `final` is set: y[g] = alpha*tmp[g] + beta*y[g] + csign*(*coef)*cvec[g].
// block (32, 8): x = gene, y = tile lane.
constexpr int ADJR_TY = 32;  // tile lanes per gene: 32 x 32 threads keep ~10 partial loads per thread at C3
__global__ void __launch_bounds__(32 * ADJR_TY) adj_reduce_kernel(const double *__restrict__ partial, int64_t ntiles, int64_t n,
                                                         const double *__restrict__ mu, double *__restrict__ tmp, int final,
                                                         double alpha, double beta, double *__restrict__ y,
                                                         const double *__restrict__ coef, double csign,
                                                         const double *__restrict__ cvec, int use_p2p, P2PCtx pc,
                                                         const int64_t *__restrict__ esegptr, const double *__restrict__ esegsum,
                                                         const double *__restrict__ einv) {
    __shared__ double sh[ADJR_TY][33];
    __shared__ double shw[ADJR_TY];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * 32 + tx;
    double acc = 0.0, wacc = 0.0;
    for (int64_t t = ty; t < ntiles; t += ADJR_TY) {
        if (g < n) acc += partial[t * (n + 1) + g];
        if (tx == 0) wacc += partial[t * (n + 1) + n];
    }
    if (esegptr != nullptr && g < n) {
        // count-level operator: the gene's exception entries (side matrix; exact value times sd, dotted with w segment by
        // segment by adj_exceptions_kernel), scaled by 1/sd like the tile partials — fixed order, deterministic
        double e = 0.0;
        for (int64_t k = esegptr[g] + ty, e1 = esegptr[g + 1]; k < e1; k += ADJR_TY) e += esegsum[k];
        acc = fma(e, einv[g], acc);
    }
    sh[ty][tx] = acc;
    if (tx == 0) shw[ty] = wacc;
    __syncthreads();
    if (ty == 0 && g < n) {
        double s = 0.0, ws = 0.0;
#pragma unroll
        for (int i = 0; i < ADJR_TY; ++i) { s += sh[i][tx]; ws += shw[i]; }
        double v = mu ? fma(-ws, mu[g], s) : s;
        if (final) {
            double r = alpha * v;
            if (beta != 0.0) r = fma(beta, y[g], r);
            if (coef != nullptr) r = fma(csign * (*coef), cvec[g], r);
            y[g] = r;
        } else if (use_p2p) {
            p2p_store(pc, (int)g, v);  // fused exchange: the partial goes straight into every rank's mailbox (NVLink stores)
        } else {
            tmp[g] = v;
        }
    }
    if (use_p2p) p2p_publish_last_block(pc, gridDim.x);
}

// consumer of the fused exchange: wait for every rank's S'w partial, sum in rank order, apply the epilogue
// y = alpha*sum + beta*y + csign*(*coef)*cvec
__global__ void __launch_bounds__(256) p2p_combine_kernel(int64_t L, double alpha, double beta, double *__restrict__ y,
                                                          const double *__restrict__ coef, double csign,
                                                          const double *__restrict__ cvec, P2PCtx pc) {
    p2p_wait(pc);
    const double c = coef ? csign * (*coef) : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) {
        double r = alpha * p2p_sum(pc, (int)i);
        if (beta != 0.0) r = fma(beta, y[i], r);
        if (coef) r = fma(c, cvec[i], r);
        y[i] = r;
    }
}

// y = alpha*tmp + beta*y + csign*(*coef)*cvec + shift_sign*(*shift)*shiftvec_or_1
__global__ void combine_kernel(int64_t L, double alpha, const double *__restrict__ tmp, double beta, double *__restrict__ y,
                               const double *__restrict__ coef, double csign, const double *__restrict__ cvec,
                               const double *__restrict__ shift, double shift_scale, const double *__restrict__ shiftvec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    const double c = coef ? csign * (*coef) : 0.0;
    const double sh = shift ? shift_scale * (*shift) : 0.0;
    for (; i < L; i += s) {
        double v = tmp[i];
        if (shift) v = fma(sh, shiftvec ? shiftvec[i] : 1.0, v);
        double r = alpha * v;
        if (beta != 0.0) r = fma(beta, y[i], r);
        if (coef) r = fma(c, cvec[i], r);
        y[i] = r;
    }
}

// out[0] = dot(a, b) (b == null: sum(a)); single CTA, deterministic
__global__ void __launch_bounds__(1024) dot_small_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t L,
                                                         double *out) {
    __shared__ double red[32];
    double p = 0.0;
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) p = b ? fma(a[i], b[i], p) : p + a[i];
    const double t = block_sum(p, red);
    if (threadIdx.x == 0) out[0] = t;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
// one resident wave: SMs x (CTAs that fit per SM for this kernel and shared-memory size)
template <typename K>
static int resident_grid(K kernel, size_t smem, int threads = 256) {
    int per_sm = 0;
    SVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    return std::max(1, per_sm) * ctx().sm_count;
}

int64_t *make_cta_ranges(const int64_t *ptr, int64_t nseg, int64_t nnz, int G) {
    int64_t *out = nullptr;
    SVB_CUDA(cudaMalloc((void **)&out, (size_t)(G + 1) * sizeof(int64_t)));
    cta_ranges_kernel<<<(G + 256) / 256, 256, 0, ctx().stream>>>(ptr, nseg, nnz, G, out);
    count_launch();
    SVB_LAUNCH_CHECK();
    return out;
}

template <typename V, typename IdxT, int LPS, bool XSMEM, int BLOCK>
static void launch_fwd_block(svb_operator_s *op, size_t smem, double alpha, const double *dx, double beta, double *dy,
                             const double *coef, double csign, const double *cvec) {
    auto k = spmv_fwd_kernel<V, IdxT, LPS, XSMEM, BLOCK>;
    if (smem > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (op->fwd_grid == 0) {
        op->fwd_grid = (int)std::min<int64_t>(resident_grid(k, smem, BLOCK), std::max<int64_t>(1, op->m / 8));
        op->fwd_ranges = make_cta_ranges(op->rowptr, op->m, op->nnz, op->fwd_grid);
    }
    k<<<(unsigned)op->fwd_grid, BLOCK, smem, ctx().stream>>>(op->rowptr, (const IdxT *)op->fidx, (const V *)op->fval, op->m, op->n,
                                                             op->nnz, dx, op->mu, alpha, beta, dy, coef, csign, cvec, op->fwd_ranges);
    SVB_LAUNCH_CHECK();
}

template <typename V, typename IdxT, int LPS>
static void launch_fwd_lps(svb_operator_s *op, double alpha, const double *dx, double beta, double *dy, const double *coef,
                           double csign, const double *cvec) {
    Context &C = ctx();
    const size_t xs_bytes = (size_t)op->n * sizeof(double);
    const bool xsmem = xs_bytes + 256 + 1024 <= C.smem_optin;
    const size_t smem = 32 * sizeof(double) + (xsmem ? xs_bytes : 0);
    // a large gene vector leaves little of the 228 KB for L1 when six 256-thread CTAs each hold a copy: share one
    // copy among 512 threads instead (three CTAs per SM keep the same number of warps)
    if (!xsmem) launch_fwd_block<V, IdxT, LPS, false, 256>(op, smem, alpha, dx, beta, dy, coef, csign, cvec);
    else if (smem > 24 * 1024) launch_fwd_block<V, IdxT, LPS, true, 512>(op, smem, alpha, dx, beta, dy, coef, csign, cvec);
    else launch_fwd_block<V, IdxT, LPS, true, 256>(op, smem, alpha, dx, beta, dy, coef, csign, cvec);
}

template <typename V, typename IdxT>
static void launch_fwd(svb_operator_s *op, double alpha, const double *dx, double beta, double *dy, const double *coef,
                       double csign, const double *cvec) {
    const double avg = op->m > 0 ? (double)op->nnz / (double)op->m : 0.0;
    int lps = op->fwd_lps;
    if (lps == 0) lps = avg >= 256 ? 32 : avg >= 96 ? 16 : avg >= 24 ? 8 : 4;
    switch (lps) {
        case 32: launch_fwd_lps<V, IdxT, 32>(op, alpha, dx, beta, dy, coef, csign, cvec); break;
        case 16: launch_fwd_lps<V, IdxT, 16>(op, alpha, dx, beta, dy, coef, csign, cvec); break;
        case 8: launch_fwd_lps<V, IdxT, 8>(op, alpha, dx, beta, dy, coef, csign, cvec); break;
        default: launch_fwd_lps<V, IdxT, 4>(op, alpha, dx, beta, dy, coef, csign, cvec); break;
    }
}

template <typename V, int LPS>
static void launch_adj_lps(svb_operator_s *op, const double *dx) {
    Context &C = ctx();
    const size_t smem = (32 + (size_t)op->R) * sizeof(double);
    auto k = spmv_adj_kernel<V, LPS>;
    if (smem > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (op->adj_grid == 0) {
        const int64_t nseg = op->ntiles * op->n;
        op->adj_grid = (int)std::max<int64_t>(1, std::min<int64_t>(resident_grid(k, smem), nseg / 8 + 1));
        op->adj_ranges = make_cta_ranges(op->gptr, nseg, op->nnz, op->adj_grid);
    }
    k<<<(unsigned)op->adj_grid, 256, smem, C.stream>>>(op->gptr, op->rloc, (const V *)op->aval, op->m, op->n, (int)op->log2R,
                                                       op->ntiles, op->nnz, dx, op->partial, op->adj_ranges);
    SVB_LAUNCH_CHECK();
}

template <typename V>
static void launch_adj(svb_operator_s *op, const double *dx) {
    const double avg = (op->n > 0 && op->ntiles > 0) ? (double)op->nnz / ((double)op->n * (double)op->ntiles) : 0.0;
    int lps = op->adj_lps;
    if (lps == 0) lps = avg >= 256 ? 32 : avg >= 96 ? 16 : avg >= 24 ? 8 : 4;
    switch (lps) {
        case 32: launch_adj_lps<V, 32>(op, dx); break;
        case 16: launch_adj_lps<V, 16>(op, dx); break;
        case 8: launch_adj_lps<V, 8>(op, dx); break;
        default: launch_adj_lps<V, 4>(op, dx); break;
    }
}

static inline unsigned grid1d(int64_t n, int threads = 256) {
    int64_t b = (n + threads - 1) / threads;
    b = std::max<int64_t>(1, std::min<int64_t>(b, 148 * 8));
    return (unsigned)b;
}

// algorithmic bytes of one product (with this build's storage widths)
static double fwd_bytes(const svb_operator_s *op) {
    if (op->dense) return 8.0 * ((double)op->m * op->n + op->m + op->n);
    return (double)op->nnz * (op->vbytes + op->ibytes) + 8.0 * (op->m + 1) + 8.0 * op->n + 8.0 * op->m;
}
static double adj_bytes(const svb_operator_s *op) {
    if (op->dense) return 8.0 * ((double)op->m * op->n + op->m + op->n);
    return (double)op->nnz * (op->vbytes + 2) + 8.0 * ((double)op->ntiles * op->n + 1) + 8.0 * op->m + 8.0 * op->n;
}

void op_apply(svb_operator_s *op, bool trans, double alpha, const double *dx, double beta, double *dy, const double *coef,
              double csign, const double *cvec) {
    Context &C = ctx();
    cudaStream_t st = C.stream;
    if (op->dense) {
        // stored array is (rows x cols) column-major; operator = stored or stored'
        const bool use_t = (trans != op->dense_transposed);
        const int64_t rows = op->dense_transposed ? op->n : op->m;  // rows of the stored array
        const int64_t cols = op->dense_transposed ? op->m : op->n;
        const int64_t outL = trans ? op->n : op->m;
        {
            KTimer kt(trans ? SVB_K_SPMV_ADJ : SVB_K_SPMV_FWD, trans ? adj_bytes(op) : fwd_bytes(op), 0);
            if (use_t) {
                ts_gemv_t(op->dA, op->lda, rows, (int)cols, dx, op->tmp, -1);
            } else {
                ts_gemv_n(op->dA, op->lda, rows, (int)cols, dx, 1.0, 0.0, op->tmp, nullptr, -1);
            }
        }
        if (trans && C.nranks > 1) comm_allreduce_dev(op->tmp, outL);
        const double *shift = nullptr, *shiftvec = nullptr;
        if (op->mu) {
            // forward: - dot(mu, x) * 1 ; adjoint: - sum(x) * mu
            dot_small_kernel<<<1, 1024, 0, st>>>(trans ? dx : op->mu, trans ? nullptr : dx, trans ? op->m : op->n, op->scal);
            count_launch();
            if (trans && C.nranks > 1) comm_allreduce_dev(op->scal, 1);
            shift = op->scal;
            shiftvec = trans ? op->mu : nullptr;
        }
        combine_kernel<<<grid1d(outL), 256, 0, st>>>(outL, alpha, op->tmp, beta, dy, coef, csign, cvec, shift, -1.0, shiftvec);
        count_launch();
        SVB_LAUNCH_CHECK();
        return;
    }
    svb_factored_s *fc = op->fact;
    if (!trans && fc) {
        KTimer kt(SVB_K_SPMV_FWD, fact_fwd_bytes(op));
        fact_fwd(op, alpha, dx, beta, dy, coef, csign, cvec);
        return;
    }
    if (!trans) {
        KTimer kt(SVB_K_SPMV_FWD, fwd_bytes(op));
        if (op->vbytes == 8) {
            if (op->ibytes == 2) launch_fwd<double, uint16_t>(op, alpha, dx, beta, dy, coef, csign, cvec);
            else launch_fwd<double, int32_t>(op, alpha, dx, beta, dy, coef, csign, cvec);
        } else {
            if (op->ibytes == 2) launch_fwd<float, uint16_t>(op, alpha, dx, beta, dy, coef, csign, cvec);
            else launch_fwd<float, int32_t>(op, alpha, dx, beta, dy, coef, csign, cvec);
        }
        return;
    }
    const bool multi = C.nranks > 1;
    P2PCtx pc{};
    static const bool allow_fused = getenv("SVB_P2P_UNFUSED") == nullptr;
    const bool fused = multi && allow_fused && p2p_next_ctx(op->n, &pc);
    {
        KTimer kt(SVB_K_SPMV_ADJ, fc ? fact_adj_bytes(op) : adj_bytes(op), 2);
        if (fc) fact_adj_stage1(op, dx);
        else if (op->vbytes == 8) launch_adj<double>(op, dx);
        else launch_adj<float>(op, dx);
        adj_reduce_kernel<<<(unsigned)((op->n + 31) / 32), 32 * ADJR_TY, 0, st>>>(fc ? fc->partial : op->partial, fc ? fc->ntiles : op->ntiles,
                                                                          op->n, op->mu, op->tmp,
                                                                          multi ? 0 : 1, alpha, beta, dy, coef, csign, cvec,
                                                                          fused ? 1 : 0, pc, (fc && fc->exc) ? fc->e_segptr : nullptr,
                                                                          (fc && fc->exc) ? fc->e_segsum : nullptr, fc ? fc->inv : nullptr);
        SVB_LAUNCH_CHECK();
    }
    if (fused) {
        KTimer kt(SVB_K_COMM, 8.0 * op->n, 1);
        p2p_combine_kernel<<<grid1d(op->n), 256, 0, st>>>(op->n, alpha, beta, dy, coef, csign, cvec, pc);
        SVB_LAUNCH_CHECK();
    } else if (multi) {
        comm_allreduce_dev(op->tmp, op->n);
        combine_kernel<<<grid1d(op->n), 256, 0, st>>>(op->n, alpha, op->tmp, beta, dy, coef, csign, cvec, nullptr, 0.0, nullptr);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
}

// ---------------------------------------------------------------------------------------------
// construction
// ---------------------------------------------------------------------------------------------
template <typename IdxT>
__global__ void narrow_index_kernel(const int32_t *in, IdxT *out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) out[i] = (IdxT)in[i];
}
template <typename VI, typename VO>
__global__ void cast_value_kernel(const VI *in, VO *out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) out[i] = (VO)in[i];
}

template <typename VI, typename VO>
static void build_sparse(svb_operator_s *op, const svb_matrix_s *a, bool transposed, int log2R) {
    cudaStream_t st = ctx().stream;
    TileCSC<VO> tc;
    DevBuf<int64_t> rowptr;
    const bool narrow = op->n <= 65535;
    op->ibytes = narrow ? 2 : 4;
    op->vbytes = (int)sizeof(VO);
    if (!transposed) {
        build_tilecsc<VI, VO>(a, log2R, tc);
        if (narrow) {
            DevBuf<uint16_t> fidx;
            DevBuf<VO> fval;
            csr_from_tilecsc<VO, uint16_t>(tc, a, rowptr, fidx, fval);
            op->fidx = fidx.take();
            op->fval = fval.take();
        } else {
            DevBuf<int32_t> fidx;
            DevBuf<VO> fval;
            csr_from_tilecsc<VO, int32_t>(tc, a, rowptr, fidx, fval);
            op->fidx = fidx.take();
            op->fval = fval.take();
        }
        op->rowptr = rowptr.take();
    } else {
        // a is (genes x cells): its columns are the cells = the CSR of S directly
        build_tilecsc_from_transposed<VI, VO>(a, log2R, tc);
        const int64_t nnz1 = std::max<int64_t>(a->nnz, 1);
        SVB_CUDA(cudaMalloc((void **)&op->rowptr, (size_t)(a->ncol + 1) * sizeof(int64_t)));
        SVB_CUDA(cudaMemcpyAsync(op->rowptr, a->colptr, (size_t)(a->ncol + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        SVB_CUDA(cudaMalloc(&op->fidx, (size_t)nnz1 * op->ibytes));
        SVB_CUDA(cudaMalloc(&op->fval, (size_t)nnz1 * sizeof(VO)));
        if (a->nnz > 0) {
            if (narrow) narrow_index_kernel<uint16_t><<<grid1d(a->nnz), 256, 0, st>>>(a->rowidx, (uint16_t *)op->fidx, a->nnz);
            else narrow_index_kernel<int32_t><<<grid1d(a->nnz), 256, 0, st>>>(a->rowidx, (int32_t *)op->fidx, a->nnz);
            cast_value_kernel<VI, VO><<<grid1d(a->nnz), 256, 0, st>>>((const VI *)a->val, (VO *)op->fval, a->nnz);
            count_launch(2);
            SVB_LAUNCH_CHECK();
        }
        SVB_CUDA(cudaStreamSynchronize(st));
    }
    op->R = tc.R;
    op->log2R = tc.log2R;
    op->ntiles = tc.ntiles;
    op->gptr = tc.gptr.take();
    op->rloc = tc.rloc.take();
    op->aval = tc.aval.take();
}

}  // namespace svb

extern "C" {

static int operator_create_impl(svb_matrix_t a, const double *mu, int transposed, int value_storage, svb_operator_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && out, SVB_EARG, "svb_operator_create: null argument");
    Context &C = ctx();
    auto *op = new svb_operator_s();
    try {
        op->m = transposed ? a->ncol : a->nrow;
        op->n = transposed ? a->nrow : a->ncol;
        op->nnz = a->nnz;
        SVB_CHECK(op->m >= 1 && op->n >= 1, SVB_EDIM, "svb_operator_create: empty operator");
        SVB_CHECK(op->m < 2147483647LL && op->n < 2147483647LL, SVB_EDIM, "svb_operator_create: dimension too large");
        int log2R = env_int("SVB_ADJ_LOG2R", 12);
        log2R = std::max(8, std::min(log2R, 14));
        while (log2R > 8 && (1ll << (log2R - 1)) >= op->m) --log2R;  // small inputs: one small tile
        SVB_CHECK(value_storage == 0 || value_storage == SVB_F32 || value_storage == SVB_F64, SVB_EARG,
                  "svb_operator_create_ex: value_storage must be 0, SVB_F32 or SVB_F64");
        const bool to_f32 = value_storage == SVB_F32;
        switch (a->vtype) {
            case SVB_F64:
                if (to_f32) build_sparse<double, float>(op, a, transposed != 0, log2R);
                else build_sparse<double, double>(op, a, transposed != 0, log2R);
                break;
            case SVB_F32: build_sparse<float, float>(op, a, transposed != 0, log2R); break;
            case SVB_I32: build_sparse<int32_t, double>(op, a, transposed != 0, log2R); break;
            default: throw Error(SVB_EARG, "bad vtype");
        }
        op->fwd_lps = env_int("SVB_FWD_LPS", 0);
        op->adj_lps = env_int("SVB_ADJ_LPS", 0);
        SVB_CUDA(cudaMalloc((void **)&op->partial, (size_t)op->ntiles * (op->n + 1) * sizeof(double)));
        SVB_CUDA(cudaMalloc((void **)&op->tmp, (size_t)std::max(op->m, op->n) * sizeof(double)));
        SVB_CUDA(cudaMalloc((void **)&op->scal, 8 * sizeof(double)));
        if (mu) {
            SVB_CUDA(cudaMalloc((void **)&op->mu, (size_t)op->n * sizeof(double)));
            SVB_CUDA(cudaMemcpyAsync(op->mu, mu, (size_t)op->n * sizeof(double), cudaMemcpyHostToDevice, C.stream));
        }
        SVB_CUDA(cudaStreamSynchronize(C.stream));
    } catch (...) {
        delete op;
        throw;
    }
    *out = op;
    SVB_API_END
}

int svb_operator_create(svb_matrix_t a, const double *mu, int transposed, svb_operator_t *out) {
    return operator_create_impl(a, mu, transposed, 0, out);
}

int svb_operator_create_ex(svb_matrix_t a, const double *mu, int transposed, int value_storage, svb_operator_t *out) {
    return operator_create_impl(a, mu, transposed, value_storage, out);
}

int svb_operator_create_dense(int64_t m, int64_t n, const double *a, int64_t lda, const double *mu, int transposed,
                              svb_operator_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && out, SVB_EARG, "svb_operator_create_dense: null argument");
    // (m x n) is the shape of the stored array; the operator is the array or its adjoint
    SVB_CHECK(m >= 1 && n >= 1 && lda >= m, SVB_EDIM, "svb_operator_create_dense: bad dimensions");
    Context &C = ctx();
    auto *op = new svb_operator_s();
    try {
        op->dense = true;
        op->dense_transposed = transposed != 0;
        op->m = transposed ? n : m;
        op->n = transposed ? m : n;
        op->nnz = m * n;
        op->lda = m;
        op->ibytes = 0;
        SVB_CUDA(cudaMalloc((void **)&op->dA, (size_t)m * n * sizeof(double)));
        SVB_CUDA(cudaMemcpy2DAsync(op->dA, (size_t)m * 8, a, (size_t)lda * 8, (size_t)m * 8, (size_t)n, cudaMemcpyHostToDevice, C.stream));
        SVB_CUDA(cudaMalloc((void **)&op->tmp, (size_t)std::max(op->m, op->n) * sizeof(double)));
        SVB_CUDA(cudaMalloc((void **)&op->scal, 8 * sizeof(double)));
        if (mu) {
            SVB_CUDA(cudaMalloc((void **)&op->mu, (size_t)op->n * sizeof(double)));
            SVB_CUDA(cudaMemcpyAsync(op->mu, mu, (size_t)op->n * sizeof(double), cudaMemcpyHostToDevice, C.stream));
        }
        SVB_CUDA(cudaStreamSynchronize(C.stream));
    } catch (...) {
        delete op;
        throw;
    }
    *out = op;
    SVB_API_END
}

int svb_operator_free(svb_operator_t op) {
    SVB_API_BEGIN
    if (op) {
        if (ctx().initialised) cudaStreamSynchronize(ctx().stream);
        delete op;
    }
    SVB_API_END
}

int svb_operator_info(svb_operator_t op, int64_t *m, int64_t *n, int64_t *nnz, int *is_dense, int *value_bytes, int *index_bytes) {
    SVB_API_BEGIN
    SVB_CHECK(op, SVB_EARG, "null operator handle");
    if (m) *m = op->m;
    if (n) *n = op->n;
    if (nnz) *nnz = op->nnz;
    if (is_dense) *is_dense = op->dense ? 1 : 0;
    if (value_bytes) *value_bytes = op->vbytes;
    if (index_bytes) *index_bytes = op->ibytes;
    SVB_API_END
}

int svb_mul_device(svb_operator_t op, char trans, double alpha, const double *dx, double beta, double *dy) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(op && dx && dy, SVB_EARG, "svb_mul_device: null argument");
    op_apply(op, trans == 'T' || trans == 't', alpha, dx, beta, dy);
    SVB_API_END
}

int svb_mul(svb_operator_t op, char trans, double alpha, const double *x, double beta, double *y, int64_t k) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(op && x && y, SVB_EARG, "svb_mul: null argument");
    SVB_CHECK(k >= 1, SVB_EDIM, "svb_mul: k must be >= 1");
    const bool t = (trans == 'T' || trans == 't');
    const int64_t inL = t ? op->m : op->n, outL = t ? op->n : op->m;
    cudaStream_t st = ctx().stream;
    if (!op->xdev) {
        SVB_CUDA(cudaMalloc((void **)&op->xdev, (size_t)std::max(op->m, op->n) * sizeof(double)));
        SVB_CUDA(cudaMalloc((void **)&op->ydev, (size_t)std::max(op->m, op->n) * sizeof(double)));
    }
    if (k > 1 && !op->dense) {
        // matrix forms (scaling.jl:259-272): SpMM kernels, 4 right-hand sides per pass over the nonzeros
        DevBuf<double> dX((size_t)inL * k), dY((size_t)outL * k);
        SVB_CUDA(cudaMemcpyAsync(dX.p, x, (size_t)inL * k * 8, cudaMemcpyHostToDevice, st));
        if (beta != 0.0) SVB_CUDA(cudaMemcpyAsync(dY.p, y, (size_t)outL * k * 8, cudaMemcpyHostToDevice, st));
        op_apply_mm(op, t, alpha, dX.p, inL, beta, dY.p, outL, k);
        SVB_CUDA(cudaMemcpyAsync(y, dY.p, (size_t)outL * k * 8, cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        return SVB_OK;
    }
    // vector form, or dense operator: one column at a time
    for (int64_t c = 0; c < k; ++c) {
        SVB_CUDA(cudaMemcpyAsync(op->xdev, x + c * inL, (size_t)inL * 8, cudaMemcpyHostToDevice, st));
        if (beta != 0.0) SVB_CUDA(cudaMemcpyAsync(op->ydev, y + c * outL, (size_t)outL * 8, cudaMemcpyHostToDevice, st));
        op_apply(op, t, alpha, op->xdev, beta, op->ydev);
        SVB_CUDA(cudaMemcpyAsync(y + c * outL, op->ydev, (size_t)outL * 8, cudaMemcpyDeviceToHost, st));
    }
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_API_END
}

}  // extern "C"

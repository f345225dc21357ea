// knn.cu — exact k-nearest neighbours of the cells in PCA space (the step after the path).
// Replaces the external `libcell.FindNeighbours{Euclidean,Cosine}{32,64}` reached from src/neighbours.jl:33-75 (`ann!`):
// same buffers (X n x d column-major with a column stride, nn_index n x k Int32, distances n x k), but the search is
// EXACT brute force instead of the reference's randomised multi-table approximation (`ntables`, `seed` have no role) —
// the answer the reference's own test compares against (test/test_nn.jl:31-38: partialsortperm of the pairwise distances).
//
// One thread owns one query cell: its coordinates sit in registers (D = d zero-padded, compile-time: knn_padded_dims), the points
// stream through shared memory in 32 KB tiles that every thread of the CTA reads at the same address (broadcast, no bank
// conflicts: one 16-byte shared load feeds two fp64 FMAs of all 32 lanes — the ratio at which the shared pipe keeps the
// fp64 pipe busy), four points at a time for instruction-level parallelism. Ranking key: |p|^2 - 2 q.p (Euclidean; |q|^2 is
// constant per query) or -q^.p^ on unit vectors (cosine); the k best (key, index) pairs of a thread live in a sorted
// private list, and a candidate is compared with the current k-th key first, so the list is touched ~k*ln(n/k) times per
// query. Ties keep the lower index (points are visited in ascending order, insertion is strict). The reported distances
// are recomputed from the coordinates of the k winners (sqrt(sum (q-p)^2), 1 - cos), not from the ranking key.
// Bound: fp64 FMA pipe (n^2 * D FMAs); traffic = (n/128) passes over the packed n x D points, served mostly by L2.
#include "svb_internal.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace svb {
namespace {

constexpr int KNN_THREADS = 128;
constexpr int KNN_KMAX = 64;
constexpr int KNN_TILE_DOUBLES = 4096;  // 32 KB of packed points per tile
constexpr bool KNN_AUTO_MMA = false;    // fp64 tensor-core variant for D <= 64 when SVB_KNN_MMA is unset
constexpr bool KNN_AUTO_Q2 = true;      // two queries per thread for D <= 32 when unset SVB_KNN_Q (see launch_knn)

__host__ __device__ constexpr int knn_tile_points(int D) { return (KNN_TILE_DOUBLES / D) & ~3; }

// P[i*D + c] = X[i + c*ldx] (* 1/|row| for cosine), zero for d <= c < D and for the padding rows n <= i < npad.
// cn[i] = |row|^2 (Euclidean) or 0 (cosine); +inf for padding rows, so that their key is never below any threshold.
__global__ void __launch_bounds__(256) knn_pack_kernel(const double *__restrict__ X, int64_t ldx, int64_t n, int64_t npad, int d, int D,
                                                       int metric, double *__restrict__ P, double *__restrict__ cn,
                                                       int *__restrict__ nonfinite) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    double *row = P + i * D;
    if (i >= n) {
        for (int c = 0; c < D; ++c) row[c] = 0.0;
        cn[i] = __longlong_as_double(0x7ff0000000000000ll);
        return;
    }
    double ss = 0.0;
    for (int c = 0; c < d; ++c) {
        const double v = X[i + (int64_t)c * ldx];
        ss = fma(v, v, ss);
    }
    // a NaN / Inf coordinate makes every ranking key of the cell NaN: no candidate would ever enter its list (round-1 advice:
    // the finalize kernel then read P[-1]). Reject the input instead: svb_knn returns SVB_EARG.
    if (!isfinite(ss)) *nonfinite = 1;
    double scale = 1.0;
    if (metric == SVB_METRIC_COSINE) {
        const double nrm = sqrt(ss);
        scale = nrm > 0.0 ? 1.0 / nrm : 0.0;
    }
    for (int c = 0; c < D; ++c) row[c] = c < d ? X[i + (int64_t)c * ldx] * scale : 0.0;
    cn[i] = metric == SVB_METRIC_COSINE ? 0.0 : ss;
}

__device__ __forceinline__ void knn_insert(double *bs, int *bi, int k, double s, int j, double &worst) {
    int pos = k - 1;
    while (pos > 0 && bs[pos - 1] > s) {  // strict: an equal key stays behind the earlier (lower-index) point
        bs[pos] = bs[pos - 1];
        bi[pos] = bi[pos - 1];
        --pos;
    }
    bs[pos] = s;
    bi[pos] = j;
    worst = bs[k - 1];
}

// Q queries per thread: one 16-byte shared load of a point then feeds 2*Q fp64 FMAs (Q = 2 when the coordinates of two
// queries fit the register file, D <= 32).
template <int D, int Q>
__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(const double *__restrict__ P, const double *__restrict__ cn, int64_t n,
                                                          int64_t npad, int k, int include_self, int *__restrict__ nbr /* [n][k] */) {
    constexpr int TP = knn_tile_points(D);
    extern __shared__ double sm[];
    double *tile = sm;            // TP x D, row-major
    double *cs = sm + TP * D;     // TP
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    int64_t qi[Q];
    double q[Q][D];
    double bs[Q][KNN_KMAX];
    int bi[Q][KNN_KMAX];
    double worst[Q];
#pragma unroll
    for (int s = 0; s < Q; ++s) {
        qi[s] = ((int64_t)blockIdx.x * Q + s) * KNN_THREADS + threadIdx.x;
        const int64_t qrow = qi[s] < n ? qi[s] : n - 1;
#pragma unroll
        for (int i = 0; i < D; ++i) q[s][i] = P[qrow * D + i];
        for (int r = 0; r < k; ++r) {
            bs[s][r] = inf;
            bi[s][r] = -1;
        }
        if (include_self) {  // neighbours.jl `include_self`: the cell itself is its first neighbour (distance 0)
            bs[s][0] = -inf;
            bi[s][0] = (int)qrow;
        }
        worst[s] = bs[s][k - 1];
    }
    for (int64_t p0 = 0; p0 < npad; p0 += TP) {
        __syncthreads();
        {
            const double2 *src = reinterpret_cast<const double2 *>(P + p0 * D);
            double2 *dst = reinterpret_cast<double2 *>(tile);
            for (int t = threadIdx.x; t < TP * D / 2; t += KNN_THREADS) dst[t] = src[t];
            for (int t = threadIdx.x; t < TP; t += KNN_THREADS) cs[t] = cn[p0 + t];
        }
        __syncthreads();
        for (int j = 0; j < TP; j += 4) {
            const double *t0 = tile + j * D;
            double a[Q][4];
#pragma unroll
            for (int s = 0; s < Q; ++s)
#pragma unroll
                for (int u = 0; u < 4; ++u) a[s][u] = 0.0;
#pragma unroll
            for (int i = 0; i < D; i += 2) {
                double2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const double2 *>(t0 + u * D + i);
#pragma unroll
                for (int s = 0; s < Q; ++s)
#pragma unroll
                    for (int u = 0; u < 4; ++u) a[s][u] = fma(q[s][i], v[u].x, a[s][u]);
#pragma unroll
                for (int s = 0; s < Q; ++s)
#pragma unroll
                    for (int u = 0; u < 4; ++u) a[s][u] = fma(q[s][i + 1], v[u].y, a[s][u]);
            }
            const int64_t pj = p0 + j;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double c = cs[j + u];
#pragma unroll
                for (int s = 0; s < Q; ++s) {
                    const double key = fma(-2.0, a[s][u], c);
                    if (key < worst[s] && pj + u != qi[s]) knn_insert(bs[s], bi[s], k, key, (int)(pj + u), worst[s]);
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < Q; ++s)
        if (qi[s] < n)
            for (int r = 0; r < k; ++r) nbr[qi[s] * k + r] = bi[s][r];
}

// ---- fp64 tensor-core variant (mma.sync.m8n8k4 DMMA; tcgen05 has no fp64) -------------------------------------------
// ncu on the kernel above (profiles/r03_gram_knn.md): a broadcast 16-byte shared load costs two LSU wavefronts and feeds two
// warp-wide fp64 FMAs, so with one query per thread the shared pipe saturates (86 %) at 44 % of the fp64 pipe. Here a warp
// owns 32 queries as four 8 x 4 A-fragments per 4 coordinates (D doubles per lane, loaded once), streams the points as
// 4 x 8 B-fragments (one 8-byte shared load per lane per 4 DMMAs: 512 FMAs per wavefront instead of 32) and gets 8 x 8 blocks of
// dot products. Lane (g, t) of the accumulator layout sees, for query row g of each of its four query groups, the points
// 2t, 2t+1 of every 8: it keeps a private top-k list per group over that quarter of the points, and the four lanes of a row
// merge their lists at the end with two shuffles per output (lexicographic (key, index) minimum: same tie rule).
__device__ __forceinline__ void knn_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__host__ __device__ constexpr int knn_mma_stride(int D) { return (D % 16 == 0) ? D + 4 : D + 12; }  // == 4 (mod 16): conflict-free B loads
__host__ __device__ constexpr int knn_mma_tile_points(int D) { return (KNN_TILE_DOUBLES / knn_mma_stride(D)) & ~15; }

template <int D>
__global__ void __launch_bounds__(KNN_THREADS) knn_mma_kernel(const double *__restrict__ P, const double *__restrict__ cn, int64_t n,
                                                              int64_t npad, int k, int include_self, int *__restrict__ nbr /* [n][k] */) {
    constexpr int DS = knn_mma_stride(D);
    constexpr int TP = knn_mma_tile_points(D);
    constexpr int KS = D / 4;
    extern __shared__ double sm[];
    double *tile = sm;             // TP x DS
    double *cs = sm + TP * DS;     // TP
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t qb = ((int64_t)blockIdx.x * (KNN_THREADS / 32) + warp) * 32;
    int64_t qi[4];
    double aq[4][KS];
    double bs[4][KNN_KMAX];
    int bi[4][KNN_KMAX];
    double worst[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        qi[mt] = qb + mt * 8 + g;
        const int64_t qrow = qi[mt] < n ? qi[mt] : n - 1;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) aq[mt][ks] = P[qrow * D + 4 * ks + t];
        for (int r = 0; r < k; ++r) {
            bs[mt][r] = inf;
            bi[mt][r] = -1;
        }
        if (include_self && t == 0) {  // the cell itself, once among the four lanes of its row
            bs[mt][0] = -inf;
            bi[mt][0] = (int)qrow;
        }
        worst[mt] = bs[mt][k - 1];
    }
    for (int64_t p0 = 0; p0 < npad; p0 += TP) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < TP * (D / 2); idx += KNN_THREADS) {
            const int row = idx / (D / 2), c2 = idx - row * (D / 2);
            reinterpret_cast<double2 *>(tile + row * DS)[c2] = reinterpret_cast<const double2 *>(P + (p0 + row) * D)[c2];
        }
        for (int idx = threadIdx.x; idx < TP; idx += KNN_THREADS) cs[idx] = cn[p0 + idx];
        __syncthreads();
        for (int j = 0; j < TP; j += 16) {
            double acc[4][2][2];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) acc[mt][nb][0] = acc[mt][nb][1] = 0.0;
            const double *b0 = tile + (j + g) * DS + t, *b1 = b0 + 8 * DS;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double v0 = b0[4 * ks], v1 = b1[4 * ks];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    knn_dmma(acc[mt][0][0], acc[mt][0][1], aq[mt][ks], v0);
                    knn_dmma(acc[mt][1][0], acc[mt][1][1], aq[mt][ks], v1);
                }
            }
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
                const int col = j + nb * 8 + 2 * t;
                const double c0 = cs[col], c1 = cs[col + 1];
                const int64_t pj = p0 + col;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    const double k0 = fma(-2.0, acc[mt][nb][0], c0), k1 = fma(-2.0, acc[mt][nb][1], c1);
                    if (k0 < worst[mt] && pj != qi[mt]) knn_insert(bs[mt], bi[mt], k, k0, (int)pj, worst[mt]);
                    if (k1 < worst[mt] && pj + 1 != qi[mt]) knn_insert(bs[mt], bi[mt], k, k1, (int)(pj + 1), worst[mt]);
                }
            }
        }
    }
    // merge the four partial lists of every query row (lanes t = 0..3 of the row are adjacent lanes)
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        int head = 0;
        for (int r = 0; r < k; ++r) {
            double key = head < k ? bs[mt][head] : inf;
            int idx = head < k ? bi[mt][head] : 0x7fffffff;
            if (idx < 0) idx = 0x7fffffff;  // unfilled slot: loses every tie
            double mk = key;
            int mi = idx;
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const double ok = __shfl_xor_sync(0xffffffffu, mk, o);
                const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
                if (ok < mk || (ok == mk && oi < mi)) {
                    mk = ok;
                    mi = oi;
                }
            }
            if (mk == key && mi == idx) ++head;  // this lane's head won (indices are unique across the four lists)
            if (t == 0 && qi[mt] < n) nbr[qi[mt] * k + r] = mi;
        }
    }
}

template <int D>
void launch_knn_mma(const double *P, const double *cn, int64_t n, int64_t npad, int k, int include_self, int *nbr) {
    constexpr int TP = knn_mma_tile_points(D);
    const size_t smem = (size_t)(TP * knn_mma_stride(D) + TP) * sizeof(double);
    knn_mma_kernel<D><<<(unsigned)((n + KNN_THREADS - 1) / KNN_THREADS), KNN_THREADS, smem, ctx().stream>>>(P, cn, n, npad, k, include_self, nbr);
    count_launch();
    SVB_LAUNCH_CHECK();
}

// exact distances of the k winners from the packed coordinates; outputs column-major n x k like the reference's buffers
__global__ void __launch_bounds__(256) knn_finalize_kernel(const double *__restrict__ P, int D, int64_t n, int k, int metric,
                                                           const int *__restrict__ nbr, int index_base, int32_t *__restrict__ out_idx,
                                                           double *__restrict__ out_dist) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const int64_t i = t / k;
    const int r = (int)(t - i * k);
    const int j = max(nbr[t], 0);  // (slots are always filled for finite input with k <= n; never index P[-1])
    const double *a = P + i * D, *b = P + (int64_t)j * D;
    double dist;
    if (j == i) {
        dist = 0.0;  // the cell itself (include_self): 1 - q^.q^ is only 0 to rounding
    } else if (metric == SVB_METRIC_COSINE) {
        double dot = 0.0;
        for (int c = 0; c < D; ++c) dot = fma(a[c], b[c], dot);
        dist = fmax(1.0 - dot, 0.0);  // Distances.CosineDist clamps at 0
    } else {
        double ss = 0.0;
        for (int c = 0; c < D; ++c) {
            const double df = a[c] - b[c];
            ss = fma(df, df, ss);
        }
        dist = sqrt(ss);
    }
    out_idx[i + (int64_t)r * n] = (int32_t)(j + index_base);
    out_dist[i + (int64_t)r * n] = dist;
}

__global__ void __launch_bounds__(256) knn_coords_kernel(const double *__restrict__ U, const double *__restrict__ s, int64_t m, int64_t dims,
                                                         double *__restrict__ Z) {
    const int64_t total = m * dims;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        Z[i] = U[i] * s[i / m];
}

template <int D, int Q>
void launch_knn_q(const double *P, const double *cn, int64_t n, int64_t npad, int k, int include_self, int *nbr) {
    constexpr int TP = knn_tile_points(D);
    const size_t smem = (size_t)(TP * D + TP) * sizeof(double);
    const int64_t per_cta = (int64_t)KNN_THREADS * Q;
    knn_kernel<D, Q><<<(unsigned)((n + per_cta - 1) / per_cta), KNN_THREADS, smem, ctx().stream>>>(P, cn, n, npad, k, include_self, nbr);
    count_launch();
    SVB_LAUNCH_CHECK();
}

template <int D>
void launch_knn(const double *P, const double *cn, int64_t n, int64_t npad, int k, int include_self, int *nbr) {
    const int q_env = getenv("SVB_KNN_Q") ? atoi(getenv("SVB_KNN_Q")) : 0;  // 1 / 2 force the variant, 0 = choose
    if constexpr (D <= 32) {
        // two queries per thread need enough queries to fill the machine with half as many threads
        const bool auto_q2 = KNN_AUTO_Q2 && n >= (int64_t)ctx().sm_count * KNN_THREADS * 4;
        if (q_env == 2 || (q_env == 0 && auto_q2)) {
            launch_knn_q<D, 2>(P, cn, n, npad, k, include_self, nbr);
            return;
        }
    }
    launch_knn_q<D, 1>(P, cn, n, npad, k, include_self, nbr);
}

// Compile-time coordinate count the search runs with (zero-padded). Multiples of 8 in general; the widths the pipeline
// actually uses get their own even-sized instance so that no FMA is spent on padding: dims = 1:10 (docs/src/pbmc.md:157)
// runs with D = 10 instead of 16, 20 PCs with 20 instead of 24, 50 PCs with 50 instead of 56. The tensor-core variant
// needs multiples of 8 (k-steps of 4, two per B fragment pair).
constexpr bool KNN_EXACT_WIDTHS = true;
int knn_padded_dims(int d, bool mma) {
    const bool exact = getenv("SVB_KNN_WIDTHS") ? atoi(getenv("SVB_KNN_WIDTHS")) != 0 : KNN_EXACT_WIDTHS;  // 0: multiples of 8 only
    if (exact && !mma) {
        if (d == 9 || d == 10) return 10;
        if (d == 11 || d == 12) return 12;
        if (d >= 17 && d <= 20) return 20;
        if (d == 49 || d == 50) return 50;
    }
    if (d <= 64) return (d + 7) & ~7;
    return d <= 96 ? 96 : 128;
}

// Xd: device, column-major n x d with leading dimension ldx. h_idx / h_dist: host, column-major n x k.
void knn_device(const double *Xd, int64_t ldx, int64_t n, int d, int k, int metric, int include_self, int index_base, int32_t *h_idx,
                void *h_dist, int dist_type) {
    Context &C = ctx();
    cudaStream_t st = C.stream;
    SVB_CHECK(metric == SVB_METRIC_EUCLIDEAN || metric == SVB_METRIC_COSINE, SVB_EARG, "knn: unknown metric");
    SVB_CHECK(n >= 1 && n <= 0x7fffffffll && d >= 1 && ldx >= n, SVB_EDIM, "knn: bad dimensions");
    SVB_CHECK(d <= 128, SVB_EDIM, "knn: at most 128 coordinates per cell");
    SVB_CHECK(k >= 1 && k <= KNN_KMAX, SVB_EDIM, "knn: k must satisfy 1 <= k <= 64");
    SVB_CHECK(k <= n - (include_self ? 0 : 1), SVB_EDIM, "knn: fewer than k candidate neighbours");
    SVB_CHECK(dist_type == SVB_F64 || dist_type == SVB_F32, SVB_EARG, "knn: distances must be Float32 or Float64");
    const int mma_env = getenv("SVB_KNN_MMA") ? atoi(getenv("SVB_KNN_MMA")) : -1;  // 1 / 0 force the variant, unset = choose
    const bool use_mma = d <= 64 && (mma_env == 1 || (mma_env < 0 && KNN_AUTO_MMA));
    const int D = knn_padded_dims(d, use_mma);
    const int TP = use_mma ? knn_mma_tile_points(D) : knn_tile_points(D);
    const int64_t npad = (n + TP - 1) / TP * TP;
    DevBuf<double> P((size_t)npad * D), cn((size_t)npad), dist((size_t)n * k);
    DevBuf<int> nbr((size_t)n * k);
    DevBuf<int32_t> oidx((size_t)n * k);
    DevBuf<int> nonfinite(1);
    SVB_CUDA(cudaMemsetAsync(nonfinite.p, 0, sizeof(int), st));
    knn_pack_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, st>>>(Xd, ldx, n, npad, d, D, metric, P.p, cn.p, nonfinite.p);
    count_launch();
    SVB_LAUNCH_CHECK();
    {
        int bad = 0;
        SVB_CUDA(cudaMemcpyAsync(&bad, nonfinite.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        SVB_CHECK(!bad, SVB_EARG, "svb_knn: the coordinates contain NaN or Inf");
    }
    {
    KTimer kt(SVB_K_VECTOR, 8.0 * (double)npad * D * (double)((n + KNN_THREADS - 1) / KNN_THREADS), 0);  // tile bytes read by all CTAs (L2 mostly)
    if (use_mma) {
        switch (D) {
            case 8: launch_knn_mma<8>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
            case 16: launch_knn_mma<16>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
            case 24: launch_knn_mma<24>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
            case 32: launch_knn_mma<32>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
            case 40: launch_knn_mma<40>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
            case 48: launch_knn_mma<48>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
            case 56: launch_knn_mma<56>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
            default: launch_knn_mma<64>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        }
    } else
    switch (D) {
        case 8: launch_knn<8>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 10: launch_knn<10>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 12: launch_knn<12>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 16: launch_knn<16>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 20: launch_knn<20>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 24: launch_knn<24>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 32: launch_knn<32>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 40: launch_knn<40>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 48: launch_knn<48>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 50: launch_knn<50>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 56: launch_knn<56>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 64: launch_knn<64>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 96: launch_knn<96>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        default: launch_knn<128>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
    }
    }
    knn_finalize_kernel<<<(unsigned)((n * k + 255) / 256), 256, 0, st>>>(P.p, D, n, k, metric, nbr.p, index_base, oidx.p, dist.p);
    count_launch();
    SVB_LAUNCH_CHECK();
    SVB_CUDA(cudaMemcpyAsync(h_idx, oidx.p, (size_t)n * k * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (dist_type == SVB_F64) {
        SVB_CUDA(cudaMemcpyAsync(h_dist, dist.p, (size_t)n * k * sizeof(double), cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
    } else {
        std::vector<double> tmp((size_t)n * k);
        SVB_CUDA(cudaMemcpyAsync(tmp.data(), dist.p, (size_t)n * k * sizeof(double), cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        float *o = static_cast<float *>(h_dist);
        for (size_t i = 0; i < tmp.size(); ++i) o[i] = (float)tmp[i];
    }
}

}  // namespace
}  // namespace svb

using namespace svb;

extern "C" {

int svb_knn(const void *X, int dtype, int64_t n, int64_t d, int64_t ldx, int64_t k, int metric, int include_self, int index_base,
            int32_t *nn_index, void *distances) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(X && nn_index && distances, SVB_EARG, "svb_knn: null argument");
    SVB_CHECK(dtype == SVB_F64 || dtype == SVB_F32, SVB_EARG, "svb_knn: X must be Float32 or Float64");
    SVB_CHECK(n >= 1 && d >= 1 && ldx >= n, SVB_EDIM, "svb_knn: bad dimensions");
    cudaStream_t st = ctx().stream;
    DevBuf<double> Xd((size_t)n * d);
    if (dtype == SVB_F64) {
        SVB_CUDA(cudaMemcpy2DAsync(Xd.p, (size_t)n * 8, X, (size_t)ldx * 8, (size_t)n * 8, (size_t)d, cudaMemcpyHostToDevice, st));
        SVB_CUDA(cudaStreamSynchronize(st));
    } else {
        // Float32 coordinates (test/test_nn.jl:79-96): widened on the way in, distances returned as Float32
        std::vector<double> wide((size_t)n * d);
        const float *xf = static_cast<const float *>(X);
        for (int64_t c = 0; c < d; ++c)
            for (int64_t i = 0; i < n; ++i) wide[(size_t)c * n + i] = (double)xf[(size_t)c * ldx + i];
        SVB_CUDA(cudaMemcpyAsync(Xd.p, wide.data(), (size_t)n * d * 8, cudaMemcpyHostToDevice, st));
        SVB_CUDA(cudaStreamSynchronize(st));
    }
    knn_device(Xd.p, n, n, (int)std::min<int64_t>(d, 1 << 20), (int)std::min<int64_t>(k, 1 << 20), metric, include_self, index_base, nn_index,
               distances, dtype);
    SVB_API_END
}

int svb_knn_result(svb_result_t r, int64_t dims, int64_t k, int metric, int include_self, int index_base, int32_t *nn_index,
                   double *distances) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(r && nn_index && distances, SVB_EARG, "svb_knn_result: null argument");
    SVB_CHECK(ctx().nranks == 1, SVB_EARG, "svb_knn_result: cell-sharded coordinates are not supported yet (gather U first)");
    if (dims <= 0) dims = r->nu;
    SVB_CHECK(dims <= r->nu, SVB_EDIM, "svb_knn_result: dims exceeds the number of components");
    DevBuf<double> Z((size_t)r->m * dims);
    knn_coords_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((r->m * dims + 255) / 256, 148 * 8)), 256, 0, ctx().stream>>>(
        r->U, r->s, r->m, dims, Z.p);
    count_launch();
    SVB_LAUNCH_CHECK();
    knn_device(Z.p, r->m, r->m, (int)dims, (int)std::min<int64_t>(k, 1 << 20), metric, include_self, index_base, nn_index, distances, SVB_F64);
    SVB_API_END
}

}  // extern "C"

// knn.cu — exact k-nearest neighbours of the cells in PCA space (the step after the path).
// Replaces the external `libcell.FindNeighbours{Euclidean,Cosine}{32,64}` reached from src/neighbours.jl:33-75 (`ann!`):
// same buffers (X n x d column-major with a column stride, nn_index n x k Int32, distances n x k), but the search is
// EXACT brute force instead of the reference's randomised multi-table approximation (`ntables`, `seed` have no role) —
// the answer the reference's own test compares against (test/test_nn.jl:31-38: partialsortperm of the pairwise distances).
//
// One thread owns one query cell: its coordinates sit in registers (D = d rounded up to 8, compile-time), the points
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

__host__ __device__ constexpr int knn_tile_points(int D) { return (KNN_TILE_DOUBLES / D) & ~3; }

// P[i*D + c] = X[i + c*ldx] (* 1/|row| for cosine), zero for d <= c < D and for the padding rows n <= i < npad.
// cn[i] = |row|^2 (Euclidean) or 0 (cosine); +inf for padding rows, so that their key is never below any threshold.
__global__ void __launch_bounds__(256) knn_pack_kernel(const double *__restrict__ X, int64_t ldx, int64_t n, int64_t npad, int d, int D,
                                                       int metric, double *__restrict__ P, double *__restrict__ cn) {
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

template <int D>
__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(const double *__restrict__ P, const double *__restrict__ cn, int64_t n,
                                                          int64_t npad, int k, int include_self, int *__restrict__ nbr /* [n][k] */) {
    constexpr int TP = knn_tile_points(D);
    extern __shared__ double sm[];
    double *tile = sm;            // TP x D, row-major
    double *cs = sm + TP * D;     // TP
    const int64_t qi = (int64_t)blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool active = qi < n;
    const int64_t qrow = active ? qi : n - 1;
    double q[D];
#pragma unroll
    for (int i = 0; i < D; ++i) q[i] = P[qrow * D + i];
    double bs[KNN_KMAX];
    int bi[KNN_KMAX];
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (int r = 0; r < k; ++r) {
        bs[r] = inf;
        bi[r] = -1;
    }
    if (include_self) {  // neighbours.jl `include_self`: the cell itself is its first neighbour (distance 0)
        bs[0] = -inf;
        bi[0] = (int)qrow;
    }
    double worst = bs[k - 1];
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
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int i = 0; i < D; i += 2) {
                const double2 v0 = *reinterpret_cast<const double2 *>(t0 + i);
                const double2 v1 = *reinterpret_cast<const double2 *>(t0 + D + i);
                const double2 v2 = *reinterpret_cast<const double2 *>(t0 + 2 * D + i);
                const double2 v3 = *reinterpret_cast<const double2 *>(t0 + 3 * D + i);
                a0 = fma(q[i], v0.x, a0);
                a1 = fma(q[i], v1.x, a1);
                a2 = fma(q[i], v2.x, a2);
                a3 = fma(q[i], v3.x, a3);
                a0 = fma(q[i + 1], v0.y, a0);
                a1 = fma(q[i + 1], v1.y, a1);
                a2 = fma(q[i + 1], v2.y, a2);
                a3 = fma(q[i + 1], v3.y, a3);
            }
            const double s0 = fma(-2.0, a0, cs[j]), s1 = fma(-2.0, a1, cs[j + 1]);
            const double s2 = fma(-2.0, a2, cs[j + 2]), s3 = fma(-2.0, a3, cs[j + 3]);
            const int64_t pj = p0 + j;
            if (s0 < worst && pj != qi) knn_insert(bs, bi, k, s0, (int)pj, worst);
            if (s1 < worst && pj + 1 != qi) knn_insert(bs, bi, k, s1, (int)(pj + 1), worst);
            if (s2 < worst && pj + 2 != qi) knn_insert(bs, bi, k, s2, (int)(pj + 2), worst);
            if (s3 < worst && pj + 3 != qi) knn_insert(bs, bi, k, s3, (int)(pj + 3), worst);
        }
    }
    if (active)
        for (int r = 0; r < k; ++r) nbr[qi * k + r] = bi[r];
}

// exact distances of the k winners from the packed coordinates; outputs column-major n x k like the reference's buffers
__global__ void __launch_bounds__(256) knn_finalize_kernel(const double *__restrict__ P, int D, int64_t n, int k, int metric,
                                                           const int *__restrict__ nbr, int index_base, int32_t *__restrict__ out_idx,
                                                           double *__restrict__ out_dist) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const int64_t i = t / k;
    const int r = (int)(t - i * k);
    const int j = nbr[t];
    const double *a = P + i * D, *b = P + (int64_t)j * D;
    double dist;
    if (metric == SVB_METRIC_COSINE) {
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

template <int D>
void launch_knn(const double *P, const double *cn, int64_t n, int64_t npad, int k, int include_self, int *nbr) {
    constexpr int TP = knn_tile_points(D);
    const size_t smem = (size_t)(TP * D + TP) * sizeof(double);
    knn_kernel<D><<<(unsigned)((n + KNN_THREADS - 1) / KNN_THREADS), KNN_THREADS, smem, ctx().stream>>>(P, cn, n, npad, k, include_self, nbr);
    count_launch();
    SVB_LAUNCH_CHECK();
}

int knn_padded_dims(int d) {
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
    const int D = knn_padded_dims(d);
    const int TP = knn_tile_points(D);
    const int64_t npad = (n + TP - 1) / TP * TP;
    DevBuf<double> P((size_t)npad * D), cn((size_t)npad), dist((size_t)n * k);
    DevBuf<int> nbr((size_t)n * k);
    DevBuf<int32_t> oidx((size_t)n * k);
    knn_pack_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, st>>>(Xd, ldx, n, npad, d, D, metric, P.p, cn.p);
    count_launch();
    SVB_LAUNCH_CHECK();
    switch (D) {
        case 8: launch_knn<8>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 16: launch_knn<16>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 24: launch_knn<24>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 32: launch_knn<32>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 40: launch_knn<40>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 48: launch_knn<48>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 56: launch_knn<56>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 64: launch_knn<64>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        case 96: launch_knn<96>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
        default: launch_knn<128>(P.p, cn.p, n, npad, k, include_self, nbr.p); break;
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

// snn.cu — jaccard_index / shared_nearest_neighbours (src/neighbours.jl:88-131,263-270): the consumer of the kNN graph.
// The reference forms snn = nn' * nn with the generic sparse product (entry (i, j) = |N(i) ∩ N(j)|, N(i) = column i of nn =
// the neighbours of cell i), maps every stored x to x / (k + (k - x)) and drops abs(x) <= prune (`droptol!`).
//
// Here no product is materialised and nothing is sorted globally. A warp owns a column j. Every cell i that shares a
// neighbour with j is reached through the reverse lists: for p in N(j), for i in R(p) = {i : p in N(i)}. A pair (i, j) is
// reached once per common neighbour; it is EMITTED only at its smallest common neighbour — the lane that holds the
// occurrence (p, i) walks the ascending list N(i), looks every element up in the ascending list N(j) by binary search,
// stops when a common element below p turns up (a later occurrence owns the pair) and otherwise ends with the full
// intersection count: no de-duplication storage at all, integer work only. That is the GENERAL path (any neighbourhood
// sizes, hubs of any in-degree). Measured on a B200 (262,144 cells, k = 20) it is bound by the divergent 4-byte loads of
// the lists N(i) — 0.11 s — so columns whose candidate count fits take the FAST path instead: the warp counts the
// occurrences of every i in a private shared-memory hash table (one coalesced read of each reverse list, one shared
// atomic per occurrence; the count of i IS |N(i) ∩ N(j)|) and never touches N(i). Either way the survivors of the prune
// test are counted (pass 1), scanned into the output column pointers, written (pass 2) and put in ascending row order by
// rank counting inside each column (distinct keys, typically < 100 per column), so the result does not depend on the path.
// Values are computed exactly as the reference does, in the output element type: one rounded division per entry, so the
// result is bit-identical to the reference's (Float32 or Float64).
#include "svb_internal.h"

#include <algorithm>
#include <cstdlib>

using namespace svb;

namespace svb {
namespace {

constexpr bool SNN_HASH_DEFAULT = true;  // columns whose candidates fit a per-warp hash table are counted there (see below)

inline unsigned snn_grid(int64_t n, int threads = 256) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, 148 * 16));
}

// rows ascending (strictly) inside every column and inside [0, nrow): what SparseMatrixCSC guarantees and the binary
// searches below rely on. flag[0] != 0 afterwards means a malformed input.
__global__ void snn_validate_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, int64_t ncol, int64_t nrow,
                                    int *__restrict__ flag) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < ncol; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e0 = colptr[j], e1 = colptr[j + 1];
        int32_t prev = -1;
        bool bad = e1 < e0;
        for (int64_t e = e0; e < e1; ++e) {
            const int32_t r = rowidx[e];
            if (r <= prev || (int64_t)r >= nrow) bad = true;
            prev = r;
        }
        if (bad) atomicExch(flag, 1);
    }
}

// indeg[p] = #{j : p in N(j)} (exact integer atomics; the order of the additions is irrelevant)
__global__ void snn_indegree_kernel(const int32_t *__restrict__ rowidx, int64_t nnz, unsigned long long *__restrict__ indeg) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&indeg[rowidx[e]], 1ull);
}

// reverse lists: rev[rptr[p] .. rptr[p+1]) = {j : p in N(j)} in ANY order (the enumeration below does not depend on it:
// the final order of a column is fixed by the rank sort)
__global__ void snn_reverse_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, int64_t ncol,
                                   const int64_t *__restrict__ rptr, unsigned int *__restrict__ cursor, int32_t *__restrict__ rev) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < ncol; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e1 = colptr[j + 1];
        for (int64_t e = colptr[j]; e < e1; ++e) {
            const int32_t p = rowidx[e];
            const unsigned int slot = atomicAdd(&cursor[p], 1u);
            rev[rptr[p] + slot] = (int32_t)j;
        }
    }
}

// is q an element of the ascending list a[0 .. len) ?
__device__ __forceinline__ bool snn_contains(const int32_t *__restrict__ a, int len, int32_t q) {
    int lo = 0, hi = len;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < q) lo = mid + 1; else hi = mid;
    }
    return lo < len && __ldg(a + lo) == q;
}

// x / (k + (k - x)) in the output type (neighbours.jl:90: f(x) = x / (k + (k - x)); :106 the same with k = diag)
template <typename T>
__device__ __forceinline__ T snn_value(int c, T kk) {
    const T x = (T)c;
    return x / (kk + (kk - x));
}

// WRITE = false: returns the number of surviving entries of column j.
// WRITE = true : the survivors (row, intersection count) go to tmp_row / tmp_cnt at base + their position in enumeration
//                order; returns the same number. Both passes enumerate identically. Warp-collective (j is warp-uniform).
template <typename T, bool WRITE>
__device__ __forceinline__ int64_t snn_column_general(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                      const int64_t *__restrict__ rptr, const int32_t *__restrict__ rev, const int32_t *Nj,
                                                      int dj, T kk, T prune, int64_t base, int32_t *__restrict__ tmp_row,
                                                      int32_t *__restrict__ tmp_cnt, int lane) {
    int64_t total = 0;
    for (int a = 0; a < dj; ++a) {
        const int32_t p = __ldg(Nj + a);
        const int64_t r0 = rptr[p], r1 = rptr[p + 1];
        for (int64_t rb = r0; rb < r1; rb += 32) {  // warp-uniform trip count: the ballot below is collective
            const int64_t r = rb + lane;
            bool keep = false;
            int32_t i = 0;
            int c = 0;
            if (r < r1) {
                i = rev[r];
                const int64_t i0 = colptr[i], i1 = colptr[i + 1];
                bool first = true;  // p is the smallest common neighbour of i and j
                for (int64_t e = i0; e < i1; ++e) {
                    const int32_t q = __ldg(rowidx + e);
                    if (snn_contains(Nj, dj, q)) {
                        if (q < p) {
                            first = false;
                            break;
                        }
                        ++c;
                    }
                }
                if (first) keep = !(fabs((double)snn_value<T>(c, kk)) <= (double)prune);  // droptol!: abs(x) <= tol goes
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (WRITE && keep) {
                const int64_t dst = base + total + __popc(bal & ((1u << lane) - 1u));
                tmp_row[dst] = i;
                tmp_cnt[dst] = c;
            }
            total += __popc(bal);
        }
    }
    return total;
}

// Fast path: occurrences of every candidate counted in the warp's hash table (open addressing, linear probing; the caller
// guarantees candidates <= SNN_HASH_LIMIT < SNN_HASH_SLOTS, so a free slot always exists). hkey = -1 marks a free slot.
constexpr int SNN_HASH_SLOTS = 2048;                       // per warp: 2048 x (key, count) = 16 KB
constexpr int SNN_HASH_LIMIT = SNN_HASH_SLOTS * 3 / 4;     // candidate occurrences (>= distinct candidates) per column
constexpr int SNN_HASH_WARPS = 4;                          // warps per CTA on the hash path (64 KB of shared memory)

template <typename T, bool WRITE>
__device__ __forceinline__ int64_t snn_column_hash(const int64_t *__restrict__ rptr, const int32_t *__restrict__ rev, const int32_t *Nj, int dj,
                                                   T kk, T prune, int64_t base, int32_t *__restrict__ tmp_row,
                                                   int32_t *__restrict__ tmp_cnt, int lane, int *hkey, int *hcnt) {
    for (int t = lane; t < SNN_HASH_SLOTS; t += 32) {
        hkey[t] = -1;
        hcnt[t] = 0;
    }
    __syncwarp();
    for (int a = 0; a < dj; ++a) {
        const int32_t p = __ldg(Nj + a);
        const int64_t r1 = rptr[p + 1];
        for (int64_t r = rptr[p] + lane; r < r1; r += 32) {
            const int i = rev[r];
            unsigned h = ((unsigned)i * 2654435761u) >> 21;  // 11 bits
            while (true) {
                const int old = atomicCAS(&hkey[h], -1, i);
                if (old == -1 || old == i) {
                    atomicAdd(&hcnt[h], 1);
                    break;
                }
                h = (h + 1) & (SNN_HASH_SLOTS - 1);
            }
        }
    }
    __syncwarp();
    int64_t total = 0;
    for (int t0 = 0; t0 < SNN_HASH_SLOTS; t0 += 32) {
        const int i = hkey[t0 + lane];
        const int c = hcnt[t0 + lane];
        const bool keep = i >= 0 && !(fabs((double)snn_value<T>(c, kk)) <= (double)prune);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (WRITE && keep) {
            const int64_t dst = base + total + __popc(bal & ((1u << lane) - 1u));
            tmp_row[dst] = i;
            tmp_cnt[dst] = c;
        }
        total += __popc(bal);
    }
    __syncwarp();  // the table is cleared again for the warp's next column
    return total;
}

// WRITE = false: cnt[j] = number of surviving entries of column j.
// WRITE = true : cnt = exclusive scan of those counts; survivors written at cnt[j] + position.
// HASH: dynamic shared memory = (blockDim.x / 32) tables; columns with more than SNN_HASH_LIMIT candidate occurrences
// (hubs) take the general path inside the same kernel.
template <typename T, bool WRITE, bool HASH>
__global__ void __launch_bounds__(256) snn_enumerate_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                            int64_t n, const int64_t *__restrict__ rptr, const int32_t *__restrict__ rev,
                                                            int64_t kfixed, T prune, int64_t *__restrict__ cnt,
                                                            int32_t *__restrict__ tmp_row, int32_t *__restrict__ tmp_cnt) {
    extern __shared__ int snn_tables[];
    const int lane = threadIdx.x & 31;
    int *hkey = snn_tables + (threadIdx.x >> 5) * 2 * SNN_HASH_SLOTS;
    int *hcnt = hkey + SNN_HASH_SLOTS;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += nwarps) {  // j is warp-uniform
        const int64_t j0 = colptr[j];
        const int dj = (int)(colptr[j + 1] - j0);
        const int32_t *Nj = rowidx + j0;
        const T kk = (T)(kfixed > 0 ? kfixed : (int64_t)dj);
        const int64_t base = WRITE ? cnt[j] : 0;
        int64_t total;
        bool general = true;
        if (HASH) {
            // candidate occurrences of the column = sum of the in-degrees of its neighbours (warp-uniform after the reduction)
            long long cand = 0;
            for (int a = lane; a < dj; a += 32) {
                const int32_t p = __ldg(Nj + a);
                cand += rptr[p + 1] - rptr[p];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);
            general = cand > SNN_HASH_LIMIT;
        }
        if (general)
            total = snn_column_general<T, WRITE>(colptr, rowidx, rptr, rev, Nj, dj, kk, prune, base, tmp_row, tmp_cnt, lane);
        else
            total = snn_column_hash<T, WRITE>(rptr, rev, Nj, dj, kk, prune, base, tmp_row, tmp_cnt, lane, hkey, hcnt);
        if (!WRITE && lane == 0) cnt[j] = total;
    }
}

// ascending row order inside every column by rank counting (the rows of a column are distinct), value = f(count)
template <typename T>
__global__ void __launch_bounds__(256) snn_finalize_kernel(const int64_t *__restrict__ ocolptr, const int64_t *__restrict__ colptr, int64_t n,
                                                           int64_t kfixed, const int32_t *__restrict__ tmp_row,
                                                           const int32_t *__restrict__ tmp_cnt, int32_t *__restrict__ orow,
                                                           T *__restrict__ oval) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += nwarps) {
        const int64_t s0 = ocolptr[j], s1 = ocolptr[j + 1];
        const T kk = (T)(kfixed > 0 ? kfixed : (colptr[j + 1] - colptr[j]));
        for (int64_t t = s0 + lane; t < s1; t += 32) {
            const int32_t key = tmp_row[t];
            int64_t rank = 0;
            for (int64_t u = s0; u < s1; ++u) rank += (__ldg(tmp_row + u) < key) ? 1 : 0;
            orow[s0 + rank] = key;
            oval[s0 + rank] = snn_value<T>(tmp_cnt[t], kk);
        }
    }
}

template <typename T>
svb_matrix_s *jaccard_run(const svb_matrix_s *nn, int64_t k, double prune, int vtype) {
    cudaStream_t st = ctx().stream;
    const int64_t n = nn->ncol, nnz = nn->nnz;
    DevBuf<int> flag(1);
    DevBuf<int64_t> rptr((size_t)(n + 1)), cnt((size_t)(n + 1));
    DevBuf<unsigned int> cursor((size_t)std::max<int64_t>(n, 1));
    DevBuf<int32_t> rev((size_t)std::max<int64_t>(nnz, 1));
    SVB_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    SVB_CUDA(cudaMemsetAsync(rptr.p, 0, (size_t)(n + 1) * sizeof(int64_t), st));
    SVB_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)(n + 1) * sizeof(int64_t), st));
    SVB_CUDA(cudaMemsetAsync(cursor.p, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(unsigned int), st));
    if (n > 0) {
        snn_validate_kernel<<<snn_grid(n), 256, 0, st>>>(nn->colptr, nn->rowidx, n, nn->nrow, flag.p);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    int bad = 0;
    SVB_CUDA(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_CHECK(!bad, SVB_EDIM, "svb_jaccard_index: row indices must be strictly ascending inside every column and below nrow");
    // reverse neighbour lists
    if (nnz > 0) {
        snn_indegree_kernel<<<snn_grid(nnz), 256, 0, st>>>(nn->rowidx, nnz, (unsigned long long *)rptr.p);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    exclusive_scan_i64(rptr.p, n + 1, st);
    if (nnz > 0) {
        snn_reverse_kernel<<<snn_grid(n), 256, 0, st>>>(nn->colptr, nn->rowidx, n, rptr.p, cursor.p, rev.p);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    // pass 1: survivors per column -> output column pointers. SVB_SNN_HASH=0 / 1 forces the general / hash kernels.
    const int hash_env = getenv("SVB_SNN_HASH") ? atoi(getenv("SVB_SNN_HASH")) : -1;
    const bool hash = hash_env < 0 ? SNN_HASH_DEFAULT : hash_env != 0;
    const int ethreads = hash ? SNN_HASH_WARPS * 32 : 256;
    const size_t esmem = hash ? (size_t)SNN_HASH_WARPS * 2 * SNN_HASH_SLOTS * sizeof(int) : 0;
    const unsigned egrid = snn_grid(n * 32, ethreads);
    if (hash) {
        SVB_CUDA(cudaFuncSetAttribute(snn_enumerate_kernel<T, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
        SVB_CUDA(cudaFuncSetAttribute(snn_enumerate_kernel<T, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
    }
    if (nnz > 0) {
        if (hash)
            snn_enumerate_kernel<T, false, true><<<egrid, ethreads, esmem, st>>>(nn->colptr, nn->rowidx, n, rptr.p, rev.p, k, (T)prune, cnt.p, nullptr, nullptr);
        else
            snn_enumerate_kernel<T, false, false><<<egrid, ethreads, 0, st>>>(nn->colptr, nn->rowidx, n, rptr.p, rev.p, k, (T)prune, cnt.p, nullptr, nullptr);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    exclusive_scan_i64(cnt.p, n + 1, st);
    int64_t onnz = 0;
    SVB_CUDA(cudaMemcpyAsync(&onnz, cnt.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_CHECK(onnz >= 0, SVB_EDIM, "svb_jaccard_index: internal count overflow");
    svb_matrix_s *out = matrix_alloc(n, n, onnz, vtype);
    try {
        SVB_CUDA(cudaMemcpyAsync(out->colptr, cnt.p, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        if (onnz > 0) {
            DevBuf<int32_t> tmp_row((size_t)onnz), tmp_cnt((size_t)onnz);
            if (hash)
                snn_enumerate_kernel<T, true, true><<<egrid, ethreads, esmem, st>>>(nn->colptr, nn->rowidx, n, rptr.p, rev.p, k, (T)prune, cnt.p, tmp_row.p, tmp_cnt.p);
            else
                snn_enumerate_kernel<T, true, false><<<egrid, ethreads, 0, st>>>(nn->colptr, nn->rowidx, n, rptr.p, rev.p, k, (T)prune, cnt.p, tmp_row.p, tmp_cnt.p);
            snn_finalize_kernel<T><<<snn_grid(n * 32), 256, 0, st>>>(out->colptr, nn->colptr, n, k, tmp_row.p, tmp_cnt.p, out->rowidx, (T *)out->val);
            count_launch(2);
            SVB_LAUNCH_CHECK();
            SVB_CUDA(cudaStreamSynchronize(st));  // tmp_row / tmp_cnt are freed here
        }
        SVB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        delete out;
        throw;
    }
    return out;
}

}  // namespace
}  // namespace svb

extern "C" {

int svb_jaccard_index(svb_matrix_t nn, int64_t k, double prune, int dtype, svb_matrix_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(nn && out, SVB_EARG, "svb_jaccard_index: null argument");
    SVB_CHECK(nn->nrow == nn->ncol, SVB_EDIM, "svb_jaccard_index: the neighbour graph must be square (cells x cells)");
    SVB_CHECK(dtype == SVB_F32 || dtype == SVB_F64, SVB_EARG, "svb_jaccard_index: dtype must be SVB_F32 or SVB_F64");
    SVB_CHECK(k < ((int64_t)1 << 24) || dtype == SVB_F64, SVB_EDIM, "svb_jaccard_index: k is not exact in Float32");
    *out = dtype == SVB_F32 ? jaccard_run<float>(nn, k, prune, SVB_F32) : jaccard_run<double>(nn, k, prune, SVB_F64);
    SVB_API_END
}

}  // extern "C"

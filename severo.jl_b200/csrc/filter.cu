// filter.cu — filter_cells / filter_features / filter_counts (src/filtering.jl:15-35,101-106) on the device: the step
// right before the hot path. Integer work only; results are bit-exact by construction.
//   filter_cells   : features_per_cell = #{j : A_ij > min_feature_count}; keep = (features_per_cell >= min_features)
//                    and, when min_umi > 0, (sum_j A_ij > min_umi)                         (filtering.jl:22-35)
//   filter_features: cells_per_feature = #{i kept : A_ij > 0}; keep = (cells_per_feature >= min_cells)   (:15-20)
//   filter_counts  : cells first, then features on the remaining cells ("this order can be important", :84,101-103)
// The output keeps every stored entry of a kept (cell, feature) pair, explicit zeros included, rows renumbered.
#include "svb_internal.h"
#include "layout.cuh"

#include <algorithm>
#include <cstring>

using namespace svb;

namespace svb {

constexpr int FLT_SL = 64;  // contiguous slices per column (one warp each) for the ordered compaction

static inline unsigned flt_grid(int64_t n, int threads = 256) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, 148 * 16));
}

// per cell: number of features with a count above the detection threshold, and the UMI total (exact integers)
__global__ void flt_cell_stats_kernel(const int32_t *__restrict__ rowidx, const int32_t *__restrict__ val, int64_t nnz,
                                      long long min_feature_count, unsigned int *__restrict__ nfeat,
                                      unsigned long long *__restrict__ umi) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = val[i], r = rowidx[i];
        if ((long long)v > min_feature_count) atomicAdd(&nfeat[r], 1u);
        atomicAdd(&umi[r], (unsigned long long)(long long)v);
    }
}

// keep[i] (0/1) and cnt[i] = keep[i] for the scan that renumbers the kept cells
__global__ void flt_cell_keep_kernel(const unsigned int *__restrict__ nfeat, const unsigned long long *__restrict__ umi, int64_t m,
                                     long long min_features, long long min_umi, uint8_t *__restrict__ keep,
                                     int64_t *__restrict__ cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == m) cnt[m] = 0;
    if (i >= m) return;
    bool k = (long long)nfeat[i] >= min_features;
    if (min_umi > 0) k = k && ((long long)umi[i] > min_umi);  // filtering.jl:26-28 (strictly greater)
    keep[i] = k ? 1 : 0;
    cnt[i] = k ? 1 : 0;
}

// grid (ncol, FLT_SL / 8), one warp per slice: stored entries of kept cells (slicecnt) and, per column, the number of
// kept cells with a positive count (cells_per_feature)
__global__ void __launch_bounds__(256) flt_col_count_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                            const int32_t *__restrict__ val, const uint8_t *__restrict__ keep,
                                                            int64_t *__restrict__ slicecnt, unsigned long long *__restrict__ cpf) {
    const int64_t j = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int sl = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int64_t beg = colptr[j], len = colptr[j + 1] - beg;
    const int64_t k0 = beg + (len * sl) / FLT_SL, k1 = beg + (len * (sl + 1)) / FLT_SL;
    int stored = 0, pos = 0;
    for (int64_t k = k0 + lane; k < k1; k += 32) {
        const bool kept = keep[rowidx[k]] != 0;
        stored += kept;
        pos += kept && (val[k] > 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        stored += __shfl_xor_sync(0xffffffffu, stored, o);
        pos += __shfl_xor_sync(0xffffffffu, pos, o);
    }
    if (lane == 0) {
        slicecnt[j * FLT_SL + sl] = stored;
        if (pos) atomicAdd(&cpf[j], (unsigned long long)pos);
    }
}

__global__ void flt_feature_keep_kernel(const unsigned long long *__restrict__ cpf, int64_t n, long long min_cells,
                                        uint8_t *__restrict__ keep) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) keep[j] = ((long long)cpf[j] >= min_cells) ? 1 : 0;
}

// slice counts of dropped features do not take part in the output
__global__ void flt_mask_slices_kernel(int64_t *__restrict__ slicecnt, const uint8_t *__restrict__ fkeep, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * FLT_SL && !fkeep[i / FLT_SL]) slicecnt[i] = 0;
    if (i == n * FLT_SL) slicecnt[i] = 0;
}

// colptr of the output: offset of the first slice of every kept feature, in the order of the kept features
__global__ void flt_colptr_kernel(const int64_t *__restrict__ sliceoff, const int64_t *__restrict__ fmap, int64_t n, int64_t nkept,
                                  int64_t *__restrict__ ocolptr) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && fmap[j + 1] > fmap[j]) ocolptr[fmap[j]] = sliceoff[j * FLT_SL];
    if (j == n) ocolptr[nkept] = sliceoff[n * FLT_SL];
}

__global__ void __launch_bounds__(256) flt_fill_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                       const int32_t *__restrict__ val, const uint8_t *__restrict__ ckeep,
                                                       const int64_t *__restrict__ cmap, const uint8_t *__restrict__ fkeep,
                                                       const int64_t *__restrict__ sliceoff, int32_t *__restrict__ orow,
                                                       int32_t *__restrict__ oval) {
    const int64_t j = blockIdx.x;
    if (!fkeep[j]) return;
    const int lane = threadIdx.x & 31;
    const int sl = blockIdx.y * 8 + (threadIdx.x >> 5);
    int64_t pos = sliceoff[j * FLT_SL + sl];
    if (sliceoff[j * FLT_SL + sl + 1] == pos) return;
    const int64_t beg = colptr[j], len = colptr[j + 1] - beg;
    const int64_t k0 = beg + (len * sl) / FLT_SL, k1 = beg + (len * (sl + 1)) / FLT_SL;
    for (int64_t kb = k0; kb < k1; kb += 32) {
        const int64_t k = kb + lane;
        int32_t r = 0;
        bool kept = false;
        if (k < k1) {
            r = rowidx[k];
            kept = ckeep[r] != 0;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, kept);
        if (kept) {
            const int64_t dst = pos + __popc(bal & ((1u << lane) - 1u));
            orow[dst] = (int32_t)cmap[r];  // cells renumbered in their original order
            oval[dst] = val[k];
        }
        pos += __popc(bal);
    }
}

}  // namespace svb

extern "C" {

int svb_filter_counts(svb_matrix_t a, int64_t min_cells, int64_t min_features, int64_t min_feature_count, int64_t min_umi,
                      uint8_t *cell_keep, uint8_t *feature_keep, svb_matrix_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && out, SVB_EARG, "svb_filter_counts: null argument");
    SVB_CHECK(a->vtype == SVB_I32, SVB_EARG, "svb_filter_counts: integer counts required (filtering.jl works on a CountMatrix)");
    cudaStream_t st = ctx().stream;
    const int64_t m = a->nrow, n = a->ncol, nnz = a->nnz;
    DevBuf<unsigned int> nfeat((size_t)std::max<int64_t>(m, 1));
    DevBuf<unsigned long long> umi((size_t)std::max<int64_t>(m, 1)), cpf((size_t)std::max<int64_t>(n, 1));
    DevBuf<uint8_t> ckeep((size_t)std::max<int64_t>(m, 1)), fkeep((size_t)std::max<int64_t>(n, 1));
    DevBuf<int64_t> cmap((size_t)(m + 1)), fmap((size_t)(n + 1)), slices((size_t)(n * FLT_SL + 1));
    SVB_CUDA(cudaMemsetAsync(nfeat.p, 0, (size_t)std::max<int64_t>(m, 1) * sizeof(unsigned int), st));
    SVB_CUDA(cudaMemsetAsync(umi.p, 0, (size_t)std::max<int64_t>(m, 1) * sizeof(unsigned long long), st));
    SVB_CUDA(cudaMemsetAsync(cpf.p, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(unsigned long long), st));
    SVB_CUDA(cudaMemsetAsync(slices.p, 0, (size_t)(n * FLT_SL + 1) * sizeof(int64_t), st));
    if (nnz > 0) flt_cell_stats_kernel<<<flt_grid(nnz), 256, 0, st>>>(a->rowidx, (const int32_t *)a->val, nnz, min_feature_count, nfeat.p, umi.p);
    flt_cell_keep_kernel<<<(unsigned)((m + 256) / 256), 256, 0, st>>>(nfeat.p, umi.p, m, min_features, min_umi, ckeep.p, cmap.p);
    count_launch(2);
    SVB_LAUNCH_CHECK();
    exclusive_scan_i64(cmap.p, m + 1, st);
    if (n > 0 && nnz > 0) {
        dim3 grid((unsigned)n, FLT_SL / 8);
        flt_col_count_kernel<<<grid, 256, 0, st>>>(a->colptr, a->rowidx, (const int32_t *)a->val, ckeep.p, slices.p, cpf.p);
        count_launch();
    }
    flt_feature_keep_kernel<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(cpf.p, n, min_cells, fkeep.p);
    flt_mask_slices_kernel<<<(unsigned)((n * FLT_SL + 256) / 256), 256, 0, st>>>(slices.p, fkeep.p, n);
    count_launch(2);
    SVB_LAUNCH_CHECK();
    // renumber the kept features: fmap = exclusive scan of the keep flags
    {
        std::vector<uint8_t> hk((size_t)std::max<int64_t>(n, 1));
        std::vector<int64_t> hm((size_t)n + 1);
        SVB_CUDA(cudaMemcpyAsync(hk.data(), fkeep.p, (size_t)n, cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        int64_t acc = 0;
        for (int64_t j = 0; j < n; ++j) {
            hm[j] = acc;
            acc += hk[j];
        }
        hm[n] = acc;
        SVB_CUDA(cudaMemcpyAsync(fmap.p, hm.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        if (feature_keep) memcpy(feature_keep, hk.data(), (size_t)n);
        exclusive_scan_i64(slices.p, n * FLT_SL + 1, st);
        int64_t onnz = 0, mkept = 0;
        SVB_CUDA(cudaMemcpyAsync(&onnz, slices.p + n * FLT_SL, 8, cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaMemcpyAsync(&mkept, cmap.p + m, 8, cudaMemcpyDeviceToHost, st));
        if (cell_keep) SVB_CUDA(cudaMemcpyAsync(cell_keep, ckeep.p, (size_t)m, cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        svb_matrix_s *o = matrix_alloc(mkept, acc, onnz, SVB_I32);
        try {
            flt_colptr_kernel<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(slices.p, fmap.p, n, acc, o->colptr);
            count_launch();
            if (onnz > 0) {
                dim3 grid((unsigned)n, FLT_SL / 8);
                flt_fill_kernel<<<grid, 256, 0, st>>>(a->colptr, a->rowidx, (const int32_t *)a->val, ckeep.p, cmap.p, fkeep.p, slices.p,
                                                     o->rowidx, (int32_t *)o->val);
                count_launch();
            }
            SVB_LAUNCH_CHECK();
            SVB_CUDA(cudaStreamSynchronize(st));
        } catch (...) {
            delete o;
            throw;
        }
        *out = o;
    }
    SVB_API_END
}

}  // extern "C"

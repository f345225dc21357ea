// matrix.cu — device-resident CSC matrices (cells x genes, the SparseMatrixCSC layout of
// src/Severo.jl:26-34): upload / download, column subset (docs/src/pbmc.md:121), cell-range
// slice (a rank's shard), and the stable device transposes that feed the operator layouts.
#include "svb_internal.h"
#include "layout.cuh"

#include <algorithm>
#include <cstring>

using namespace svb;

svb_matrix_s::~svb_matrix_s() {
    if (colptr) cudaFree(colptr);
    if (rowidx) cudaFree(rowidx);
    if (val) cudaFree(val);
}

namespace svb {

size_t vtype_size(int vtype) {
    switch (vtype) {
        case SVB_I32: return 4;
        case SVB_I64: return 8;
        case SVB_F32: return 4;
        case SVB_F64: return 8;
    }
    throw Error(SVB_EARG, "unknown value type");
}

svb_matrix_s *matrix_alloc(int64_t nrow, int64_t ncol, int64_t nnz, int vtype) {
    SVB_CHECK(vtype == SVB_I32 || vtype == SVB_F32 || vtype == SVB_F64, SVB_EARG, "device value type must be I32/F32/F64");
    auto *a = new svb_matrix_s();
    try {
        a->nrow = nrow;
        a->ncol = ncol;
        a->nnz = nnz;
        a->vtype = vtype;
        SVB_CUDA(cudaMalloc((void **)&a->colptr, (size_t)(ncol + 1) * sizeof(int64_t)));
        if (nnz > 0) {
            SVB_CUDA(cudaMalloc((void **)&a->rowidx, (size_t)nnz * sizeof(int32_t)));
            SVB_CUDA(cudaMalloc(&a->val, (size_t)nnz * vtype_size(vtype)));
        }
    } catch (...) {
        delete a;
        throw;
    }
    return a;
}

// ---- element-wise conversion kernels ------------------------------------------------------------
template <typename TI, typename TO>
__global__ void convert_offset_kernel(const TI *in, TO *out, int64_t n, int64_t offset, int *overflow) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int64_t v = (int64_t)in[i] + offset;
        if (sizeof(TO) < 8 && (v > 2147483647LL || v < -2147483648LL) && overflow) *overflow = 1;
        out[i] = (TO)v;
    }
}

template <typename TI, typename TO>
__global__ void convert_value_kernel(const TI *in, TO *out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (TO)in[i];
}

static inline unsigned grid_for(int64_t n, int threads = 256, int max_blocks = 148 * 16) {
    int64_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (unsigned)b;
}

constexpr int64_t STAGE_ELEMS = 32ll << 20;  // staging chunk: 32 Mi elements (256 MiB as int64)

// host int (32/64, any base) -> device int32 0-based, through a staging buffer, chunked
template <typename TH>
static void upload_index(const TH *host, int64_t n, int64_t base, int32_t *dev, int *d_overflow) {
    cudaStream_t st = ctx().stream;
    const int64_t chunk = std::min<int64_t>(n, STAGE_ELEMS);
    if (n == 0) return;
    DevBuf<TH> stage((size_t)chunk);
    for (int64_t o = 0; o < n; o += chunk) {
        const int64_t c = std::min(chunk, n - o);
        SVB_CUDA(cudaMemcpyAsync(stage.p, host + o, (size_t)c * sizeof(TH), cudaMemcpyHostToDevice, st));
        convert_offset_kernel<TH, int32_t><<<grid_for(c), 256, 0, st>>>(stage.p, dev + o, c, -base, d_overflow);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    SVB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace svb

// =================================================================================================
// layout builders (declared in layout.cuh)
// =================================================================================================
namespace svb {

// startpos[t*ncol + j] = first position in column j whose row >= t*R   (t = 0..ntiles, inclusive; R = tile_rows)
__global__ void tile_bounds_kernel(const int64_t *colptr, const int32_t *rowidx, int64_t ncol, int64_t ntiles,
                                   int64_t tile_rows, int64_t *startpos) {
    const int64_t total = (ntiles + 1) * ncol;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int64_t t = i / ncol, j = i - t * ncol;
        const int64_t target = t * tile_rows;
        int64_t lo = colptr[j], hi = colptr[j + 1];
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)rowidx[mid] < target) lo = mid + 1; else hi = mid;
        }
        startpos[i] = lo;
    }
}

void tile_bounds(const svb_matrix_s *a, int64_t tile_rows, int64_t ntiles, int64_t *startpos) {
    tile_bounds_kernel<<<grid_for((ntiles + 1) * a->ncol), 256, 0, ctx().stream>>>(a->colptr, a->rowidx, a->ncol, ntiles, tile_rows, startpos);
    count_launch();
    SVB_LAUNCH_CHECK();
}

__global__ void tile_counts_kernel(const int64_t *startpos, int64_t ncol, int64_t ntiles, int64_t *cnt) {
    const int64_t total = ntiles * ncol;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) cnt[i] = startpos[i + ncol] - startpos[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[total] = 0;
}

// grid (ncol, YCH): copy column j's entries to their (tile, gene) segments
template <typename VI, typename VO>
__global__ void tile_fill_kernel(const int64_t *colptr, const int32_t *rowidx, const VI *val, int64_t ncol,
                                 int log2R, const int64_t *startpos, const int64_t *gptr, uint16_t *rloc, VO *aval) {
    const int64_t j = blockIdx.x;
    const int64_t beg = colptr[j], end = colptr[j + 1];
    for (int64_t k = beg + (int64_t)blockIdx.y * blockDim.x + threadIdx.x; k < end; k += (int64_t)gridDim.y * blockDim.x) {
        const int64_t r = rowidx[k];
        const int64_t t = r >> log2R;
        const int64_t seg = t * ncol + j;
        const int64_t dest = gptr[seg] + (k - startpos[seg]);
        rloc[dest] = (uint16_t)(r - (t << log2R));
        aval[dest] = (VO)val[k];
    }
}

template <typename VI, typename VO>
void build_tilecsc(const svb_matrix_s *a, int log2R, TileCSC<VO> &out) {
    cudaStream_t st = ctx().stream;
    const int64_t R = 1ll << log2R;
    const int64_t ntiles = std::max<int64_t>(1, (a->nrow + R - 1) / R);
    out.R = R;
    out.log2R = log2R;
    out.ntiles = ntiles;
    out.gptr.alloc((size_t)(ntiles * a->ncol + 1));
    out.rloc.alloc((size_t)std::max<int64_t>(a->nnz, 1));
    out.aval.alloc((size_t)std::max<int64_t>(a->nnz, 1));
    DevBuf<int64_t> startpos((size_t)((ntiles + 1) * a->ncol + 1));
    tile_bounds_kernel<<<grid_for((ntiles + 1) * a->ncol), 256, 0, st>>>(a->colptr, a->rowidx, a->ncol, ntiles, (int64_t)1 << log2R, startpos.p);
    tile_counts_kernel<<<grid_for(ntiles * a->ncol), 256, 0, st>>>(startpos.p, a->ncol, ntiles, out.gptr.p);
    count_launch(2);
    SVB_LAUNCH_CHECK();
    exclusive_scan_i64(out.gptr.p, ntiles * a->ncol + 1, st);
    if (a->nnz > 0 && a->ncol > 0) {
        dim3 grid((unsigned)a->ncol, 8);
        tile_fill_kernel<VI, VO><<<grid, 256, 0, st>>>(a->colptr, a->rowidx, (const VI *)a->val, a->ncol, log2R,
                                                       startpos.p, out.gptr.p, out.rloc.p, out.aval.p);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    SVB_CUDA(cudaStreamSynchronize(st));
}

__global__ void row_count_kernel(const int32_t *rowidx, int64_t nnz, unsigned long long *cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) atomicAdd(&cnt[rowidx[i]], 1ull);
}

// CSR-by-cell from the tile layout, one CTA (32 warps) per block of 32 consecutive cells.
//  phase 1 (threads over genes): inside the (tile, gene) segment the cells are ascending, so the entries of
//          this 32-cell block are one short run: find its start by binary search, walk it, and record a 32-bit
//          presence mask + the run's start offset in shared memory;
//  phase 2 (one warp per cell): ballot/popc prefix sums over the genes give every entry its position inside
//          the cell's row (ascending gene index => stable), the source offset is start + popc(mask below the
//          cell); each warp writes its row with unit stride.
// The first version walked the genes of a whole 4096-cell tile sequentially with per-cell cursors: correct, but
// its scattered 2/8-byte stores over a 23 MB region made DRAM read-modify-write every sector (60 ms at C3).
constexpr int CSRB_ROWS = 32;
constexpr int CSRB_GCH = 4096;  // genes per shared-memory chunk

// One CTA walks the row blocks [b0, b1) of one tile. `cursor[j]` is the position inside the (tile, gene j) segment
// of the first entry not yet consumed: found by ONE binary search when the CTA starts, then advanced as the
// blocks go by (a per-block binary search over every gene made the 28 k-gene transposes ~50x slower than HBM).
template <typename V, typename IdxT>
__global__ void __launch_bounds__(1024) tile_to_csr_kernel(const int64_t *__restrict__ gptr, const uint16_t *__restrict__ rloc,
                                                           const V *__restrict__ aval, int64_t ncol, int64_t nrow, int log2R,
                                                           int split, const int64_t *__restrict__ rowptr,
                                                           int64_t *__restrict__ cursor_all, IdxT *__restrict__ fidx,
                                                           V *__restrict__ fval) {
    __shared__ uint32_t mask[CSRB_GCH];
    __shared__ int64_t aoff[CSRB_GCH];
    const int64_t t = blockIdx.x / split;
    const int part = (int)(blockIdx.x - t * split);
    const int64_t R = (int64_t)1 << log2R;
    const int64_t tile_row0 = t << log2R;
    const int64_t tile_rows = min(R, nrow - tile_row0);
    const int nblocks = (int)((tile_rows + CSRB_ROWS - 1) / CSRB_ROWS);
    const int per = (nblocks + split - 1) / split;
    const int b0 = part * per, b1 = min(nblocks, b0 + per);
    if (b0 >= b1) return;
    int64_t *cursor = cursor_all + (int64_t)blockIdx.x * ncol;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t *gp = gptr + t * ncol;
    for (int blk = b0; blk < b1; ++blk) {
        const int r0l = blk * CSRB_ROWS;  // first cell of the block, local to its tile
        const int64_t myrow = tile_row0 + r0l + warp;
        int64_t pos = (myrow < nrow) ? rowptr[myrow] : 0;
        for (int64_t g0 = 0; g0 < ncol; g0 += CSRB_GCH) {
            const int ng = (int)min((int64_t)CSRB_GCH, ncol - g0);
            for (int j = threadIdx.x; j < ng; j += blockDim.x) {
                const int64_t end = __ldg(gp + g0 + j + 1);
                int64_t lo;
                if (blk == b0) {
                    lo = __ldg(gp + g0 + j);
                    int64_t hi = end;
                    while (lo < hi) {
                        const int64_t mid = (lo + hi) >> 1;
                        if ((int)__ldg(rloc + mid) < r0l) lo = mid + 1; else hi = mid;
                    }
                } else {
                    lo = cursor[g0 + j];
                }
                uint32_t mk = 0;
                int64_t k = lo;
                for (; k < end; ++k) {
                    const int rr = (int)__ldg(rloc + k) - r0l;
                    if (rr >= CSRB_ROWS) break;
                    mk |= 1u << rr;
                }
                mask[j] = mk;
                aoff[j] = lo;
                if (b1 - b0 > 1) cursor[g0 + j] = k;
            }
            __syncthreads();
            if (myrow < nrow) {
                const uint32_t below = (1u << warp) - 1u;
                for (int c = 0; c < ng; c += 32) {
                    const int j = c + lane;
                    const uint32_t mk = (j < ng) ? mask[j] : 0u;
                    const bool bit = (mk >> warp) & 1u;
                    const unsigned bal = __ballot_sync(0xffffffffu, bit);
                    if (bal == 0u) continue;
                    if (bit) {
                        const int64_t src = aoff[j] + __popc(mk & below);
                        const int64_t dst = pos + __popc(bal & ((1u << lane) - 1u));
                        fidx[dst] = (IdxT)(g0 + j);
                        fval[dst] = __ldg(aval + src);
                    }
                    pos += __popc(bal);
                }
            }
            __syncthreads();
        }
    }
}

template <typename V, typename IdxT>
void csr_from_tilecsc(const TileCSC<V> &tc, const svb_matrix_s *a, DevBuf<int64_t> &rowptr, DevBuf<IdxT> &fidx,
                      DevBuf<V> &fval) {
    cudaStream_t st = ctx().stream;
    rowptr.alloc((size_t)(a->nrow + 1));
    fidx.alloc((size_t)std::max<int64_t>(a->nnz, 1));
    fval.alloc((size_t)std::max<int64_t>(a->nnz, 1));
    SVB_CUDA(cudaMemsetAsync(rowptr.p, 0, (size_t)(a->nrow + 1) * sizeof(int64_t), st));
    if (a->nnz > 0) {
        row_count_kernel<<<grid_for(a->nnz), 256, 0, st>>>(a->rowidx, a->nnz, (unsigned long long *)rowptr.p);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    exclusive_scan_i64(rowptr.p, a->nrow + 1, st);
    if (a->nnz > 0) {
        auto kern = tile_to_csr_kernel<V, IdxT>;
        // Dense segments (HVG operator: ~9 entries per gene per 32-cell block): one CTA per block, the binary search
        // costs no more than the run itself and 40 k CTAs keep the chip full. Sparse segments (raw 28 k-gene matrices,
        // ~2 entries): tiles x split CTAs (about two resident waves) that carry per-gene cursors from block to block.
        const int blocks_per_tile = (int)(tc.R / CSRB_ROWS);
        const double run = (double)a->nnz / (std::max<double>(1.0, (double)a->ncol) * std::max<double>(1.0, (double)a->nrow / CSRB_ROWS));
        int split;
        if (run >= 4.0) {
            split = blocks_per_tile;
        } else {
            split = (int)std::max<int64_t>(1, std::min<int64_t>(blocks_per_tile, ((int64_t)ctx().sm_count * 4 + tc.ntiles - 1) / tc.ntiles));
            while (split > 1 && (int64_t)tc.ntiles * split * a->ncol * 8 > (int64_t)8 << 30) split /= 2;  // cursor memory cap: 8 GB
        }
        const bool need_cursor = split < blocks_per_tile;
        DevBuf<int64_t> cursor(need_cursor ? (size_t)tc.ntiles * split * std::max<int64_t>(a->ncol, 1) : 1);
        kern<<<(unsigned)(tc.ntiles * split), 1024, 0, st>>>(tc.gptr.p, tc.rloc.p, tc.aval.p, a->ncol, a->nrow, tc.log2R, split, rowptr.p,
                                                          cursor.p, fidx.p, fval.p);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    SVB_CUDA(cudaStreamSynchronize(st));
}

// ---- column-tile path: used when the input has many columns (cells as columns) -----------------------
// cnt[key] += 1 with key = minor * nct + coltile  (rowmajor_tiles = true, full transpose)
//            or key = coltile * nrow + minor       (rowmajor_tiles = false, tile-CSC of the transpose)
__global__ void coltile_count_kernel(const int64_t *colptr, const int32_t *rowidx, int64_t ncol, int64_t nrow,
                                     int log2R, int64_t nct, bool rowmajor_tiles, unsigned long long *cnt) {
    // one warp per column
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < ncol; c += nwarps) {
        const int64_t t = c >> log2R;
        for (int64_t k = colptr[c] + lane; k < colptr[c + 1]; k += 32) {
            const int64_t r = rowidx[k];
            const int64_t key = rowmajor_tiles ? (r * nct + t) : (t * nrow + r);
            atomicAdd(&cnt[key], 1ull);
        }
    }
}

// One CTA per column tile; columns visited in ascending order, rows distinct inside a column.
template <typename VI, typename VO, typename IdxT>
__global__ void __launch_bounds__(256) coltile_place_kernel(const int64_t *colptr, const int32_t *rowidx, const VI *val,
                                                            int64_t ncol, int64_t nrow, int log2R, int64_t nct,
                                                            bool rowmajor_tiles, const int64_t *off, int32_t *cursor,
                                                            IdxT *oidx, VO *oval, bool local_index) {
    const int64_t t = blockIdx.x;
    const int64_t c0 = t << log2R;
    const int64_t c1 = min(ncol, c0 + ((int64_t)1 << log2R));
    for (int64_t c = c0; c < c1; ++c) {
        const int64_t beg = colptr[c], end = colptr[c + 1];
        for (int64_t k = beg + threadIdx.x; k < end; k += blockDim.x) {
            const int64_t r = rowidx[k];
            const int64_t key = rowmajor_tiles ? (r * nct + t) : (t * nrow + r);
            const int32_t p = __ldcg(&cursor[key]);
            __stcg(&cursor[key], p + 1);
            const int64_t dest = off[key] + p;
            oidx[dest] = (IdxT)(local_index ? (c - c0) : c);
            oval[dest] = (VO)val[k];
        }
        __syncthreads();
    }
}

// full stable transpose through column tiles: out = a' as CSC (out.ncol = a.nrow)
template <typename V>
static svb_matrix_s *transpose_coltiles(const svb_matrix_s *a) {
    cudaStream_t st = ctx().stream;
    const int log2R = 10;
    const int64_t nct = std::max<int64_t>(1, (a->ncol + (1ll << log2R) - 1) >> log2R);
    const int64_t nkeys = a->nrow * nct;
    DevBuf<int64_t> off((size_t)(nkeys + 1));
    DevBuf<int32_t> cursor((size_t)std::max<int64_t>(nkeys, 1));
    SVB_CUDA(cudaMemsetAsync(off.p, 0, (size_t)(nkeys + 1) * sizeof(int64_t), st));
    SVB_CUDA(cudaMemsetAsync(cursor.p, 0, (size_t)std::max<int64_t>(nkeys, 1) * sizeof(int32_t), st));
    svb_matrix_s *out = matrix_alloc(a->ncol, a->nrow, a->nnz, a->vtype);
    try {
        if (a->nnz > 0) {
            coltile_count_kernel<<<grid_for(a->ncol * 32), 256, 0, st>>>(a->colptr, a->rowidx, a->ncol, a->nrow, log2R, nct,
                                                                         true, (unsigned long long *)off.p);
            count_launch();
            SVB_LAUNCH_CHECK();
        }
        exclusive_scan_i64(off.p, nkeys + 1, st);
        // out.colptr[r] = off[r * nct]
        launch_strided_copy(off.p, nct, a->nrow, out->colptr, st);
        SVB_CUDA(cudaMemcpyAsync(out->colptr + a->nrow, off.p + nkeys, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        if (a->nnz > 0) {
            coltile_place_kernel<V, V, int32_t><<<(unsigned)nct, 256, 0, st>>>(
                a->colptr, a->rowidx, (const V *)a->val, a->ncol, a->nrow, log2R, nct, true, off.p, cursor.p, out->rowidx,
                (V *)out->val, false);
            count_launch();
            SVB_LAUNCH_CHECK();
        }
        SVB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        delete out;
        throw;
    }
    return out;
}

__global__ void strided_copy_kernel(const int64_t *src, int64_t stride, int64_t n, int64_t *dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) dst[i] = src[i * stride];
}
void launch_strided_copy(const int64_t *src, int64_t stride, int64_t n, int64_t *dst, cudaStream_t st) {
    if (n <= 0) return;
    strided_copy_kernel<<<grid_for(n), 256, 0, st>>>(src, stride, n, dst);
    count_launch();
    SVB_LAUNCH_CHECK();
}

// full stable transpose through row tiles (good when a has few columns)
template <typename V>
static svb_matrix_s *transpose_rowtiles(const svb_matrix_s *a) {
    TileCSC<V> tc;
    build_tilecsc<V, V>(a, 13, tc);
    DevBuf<int64_t> rowptr;
    DevBuf<int32_t> fidx;
    DevBuf<V> fval;
    csr_from_tilecsc<V, int32_t>(tc, a, rowptr, fidx, fval);
    auto *out = new svb_matrix_s();
    out->nrow = a->ncol;
    out->ncol = a->nrow;
    out->nnz = a->nnz;
    out->vtype = a->vtype;
    out->colptr = rowptr.take();
    out->rowidx = fidx.take();
    out->val = fval.take();
    return out;
}

svb_matrix_s *matrix_transpose(const svb_matrix_s *a) {
    // row tiles (tile layout + cursor kernel) whenever the ntiles x ncol segment pointers stay below ~4 GB
    const int64_t ntiles13 = std::max<int64_t>(1, (a->nrow + 8191) / 8192);
    const bool use_rowtiles = ntiles13 * a->ncol <= ((int64_t)1 << 29);
    switch (a->vtype) {
        case SVB_I32: return use_rowtiles ? transpose_rowtiles<int32_t>(a) : transpose_coltiles<int32_t>(a);
        case SVB_F32: return use_rowtiles ? transpose_rowtiles<float>(a) : transpose_coltiles<float>(a);
        case SVB_F64: return use_rowtiles ? transpose_rowtiles<double>(a) : transpose_coltiles<double>(a);
    }
    throw Error(SVB_EARG, "bad vtype");
}

// tile-CSC of a' where a is (genes x cells) CSC: cells tiled by R, gene-major inside a tile.
template <typename VI, typename VO>
void build_tilecsc_from_transposed(const svb_matrix_s *a, int log2R, TileCSC<VO> &out) {
    cudaStream_t st = ctx().stream;
    const int64_t R = 1ll << log2R;
    const int64_t ncells = a->ncol, ngenes = a->nrow;
    const int64_t ntiles = std::max<int64_t>(1, (ncells + R - 1) / R);
    const int64_t nkeys = ntiles * ngenes;
    out.R = R;
    out.log2R = log2R;
    out.ntiles = ntiles;
    out.gptr.alloc((size_t)(nkeys + 1));
    out.rloc.alloc((size_t)std::max<int64_t>(a->nnz, 1));
    out.aval.alloc((size_t)std::max<int64_t>(a->nnz, 1));
    DevBuf<int32_t> cursor((size_t)std::max<int64_t>(nkeys, 1));
    SVB_CUDA(cudaMemsetAsync(out.gptr.p, 0, (size_t)(nkeys + 1) * sizeof(int64_t), st));
    SVB_CUDA(cudaMemsetAsync(cursor.p, 0, (size_t)std::max<int64_t>(nkeys, 1) * sizeof(int32_t), st));
    if (a->nnz > 0) {
        coltile_count_kernel<<<grid_for(ncells * 32), 256, 0, st>>>(a->colptr, a->rowidx, ncells, ngenes, log2R, ntiles, false,
                                                                    (unsigned long long *)out.gptr.p);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    exclusive_scan_i64(out.gptr.p, nkeys + 1, st);
    if (a->nnz > 0) {
        coltile_place_kernel<VI, VO, uint16_t><<<(unsigned)ntiles, 256, 0, st>>>(
            a->colptr, a->rowidx, (const VI *)a->val, ncells, ngenes, log2R, ntiles, false, out.gptr.p, cursor.p, out.rloc.p,
            out.aval.p, true);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    SVB_CUDA(cudaStreamSynchronize(st));
}

// explicit instantiations used by operator.cu
template void build_tilecsc<double, double>(const svb_matrix_s *, int, TileCSC<double> &);
template void build_tilecsc<float, float>(const svb_matrix_s *, int, TileCSC<float> &);
template void build_tilecsc<int32_t, double>(const svb_matrix_s *, int, TileCSC<double> &);
template void build_tilecsc<float, double>(const svb_matrix_s *, int, TileCSC<double> &);
template void build_tilecsc<double, float>(const svb_matrix_s *, int, TileCSC<float> &);
template void build_tilecsc<uint8_t, uint8_t>(const svb_matrix_s *, int, TileCSC<uint8_t> &);
template void csr_from_tilecsc<uint8_t, uint16_t>(const TileCSC<uint8_t> &, const svb_matrix_s *, DevBuf<int64_t> &, DevBuf<uint16_t> &, DevBuf<uint8_t> &);
template void build_tilecsc_from_transposed<double, float>(const svb_matrix_s *, int, TileCSC<float> &);
template void build_tilecsc_from_transposed<double, double>(const svb_matrix_s *, int, TileCSC<double> &);
template void build_tilecsc_from_transposed<float, float>(const svb_matrix_s *, int, TileCSC<float> &);
template void build_tilecsc_from_transposed<int32_t, double>(const svb_matrix_s *, int, TileCSC<double> &);
template void csr_from_tilecsc<double, uint16_t>(const TileCSC<double> &, const svb_matrix_s *, DevBuf<int64_t> &, DevBuf<uint16_t> &, DevBuf<double> &);
template void csr_from_tilecsc<double, int32_t>(const TileCSC<double> &, const svb_matrix_s *, DevBuf<int64_t> &, DevBuf<int32_t> &, DevBuf<double> &);
template void csr_from_tilecsc<float, uint16_t>(const TileCSC<float> &, const svb_matrix_s *, DevBuf<int64_t> &, DevBuf<uint16_t> &, DevBuf<float> &);
template void csr_from_tilecsc<float, int32_t>(const TileCSC<float> &, const svb_matrix_s *, DevBuf<int64_t> &, DevBuf<int32_t> &, DevBuf<float> &);

// ---- subset / slice ---------------------------------------------------------------------------------
__global__ void subset_len_kernel(const int64_t *colptr, const int64_t *idx, int64_t k, int64_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) out[i] = colptr[idx[i] + 1] - colptr[idx[i]];
    if (i == k) out[i] = 0;
}

template <typename V>
__global__ void subset_copy_kernel(const int64_t *colptr, const int32_t *rowidx, const V *val, const int64_t *idx,
                                   const int64_t *ocolptr, int32_t *orow, V *oval, int32_t rowshift, const int64_t *srcbeg) {
    const int64_t j = blockIdx.x;
    const int64_t src = srcbeg ? srcbeg[j] : colptr[idx[j]];
    const int64_t dst = ocolptr[j];
    const int64_t len = ocolptr[j + 1] - dst;
    for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < len; e += (int64_t)gridDim.y * blockDim.x) {
        orow[dst + e] = rowidx[src + e] - rowshift;
        oval[dst + e] = val[src + e];
    }
}

__global__ void slice_bounds_kernel(const int64_t *colptr, const int32_t *rowidx, int64_t ncol, int64_t r0, int64_t r1,
                                    int64_t *srcbeg, int64_t *len) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == ncol) len[j] = 0;
    if (j >= ncol) return;
    int64_t lo = colptr[j], hi = colptr[j + 1];
    const int64_t end = hi;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rowidx[mid] < r0) lo = mid + 1; else hi = mid;
    }
    const int64_t b = lo;
    hi = end;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rowidx[mid] < r1) lo = mid + 1; else hi = mid;
    }
    srcbeg[j] = b;
    len[j] = lo - b;
}

template <typename V>
static void launch_subset_copy(const svb_matrix_s *a, const int64_t *d_idx, svb_matrix_s *out, int32_t rowshift,
                               const int64_t *srcbeg) {
    if (out->nnz == 0 || out->ncol == 0) return;
    dim3 grid((unsigned)out->ncol, 8);
    subset_copy_kernel<V><<<grid, 256, 0, ctx().stream>>>(a->colptr, a->rowidx, (const V *)a->val, d_idx, out->colptr,
                                                          out->rowidx, (V *)out->val, rowshift, srcbeg);
    count_launch();
    SVB_LAUNCH_CHECK();
}

static svb_matrix_s *finish_subset(const svb_matrix_s *a, DevBuf<int64_t> &lens, int64_t k, int64_t nrow_out,
                                   const int64_t *d_idx, int32_t rowshift, const int64_t *srcbeg) {
    cudaStream_t st = ctx().stream;
    exclusive_scan_i64(lens.p, k + 1, st);
    int64_t nnz = 0;
    SVB_CUDA(cudaMemcpyAsync(&nnz, lens.p + k, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    svb_matrix_s *out = matrix_alloc(nrow_out, k, nnz, a->vtype);
    try {
        SVB_CUDA(cudaMemcpyAsync(out->colptr, lens.p, (size_t)(k + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        switch (a->vtype) {
            case SVB_I32: launch_subset_copy<int32_t>(a, d_idx, out, rowshift, srcbeg); break;
            case SVB_F32: launch_subset_copy<float>(a, d_idx, out, rowshift, srcbeg); break;
            case SVB_F64: launch_subset_copy<double>(a, d_idx, out, rowshift, srcbeg); break;
        }
        SVB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        delete out;
        throw;
    }
    return out;
}

}  // namespace svb

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

}  // extern "C"

// round-1 advice: a malformed matrix arriving through the public ABI must not reach the kernels that index with its rows
// (atomicAdd(&nfeat[r]) in filter.cu, the binary searches of tile_bounds, ...). One cheap pass over colptr / rowidx after the
// upload: pointers non-decreasing, rows inside [0, nrow) and strictly ascending inside a column (what SparseMatrixCSC guarantees).
__global__ void csc_validate_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, int64_t nrow, int64_t ncol,
                                    int64_t nnz, int *__restrict__ bad) {
    const int64_t j = blockIdx.x;
    if (j >= ncol) return;
    const int64_t b = colptr[j], e = colptr[j + 1];
    if (b > e || b < 0 || e > nnz) {
        if (threadIdx.x == 0) atomicOr(bad, 1);
        return;
    }
    for (int64_t k = b + threadIdx.x; k < e; k += blockDim.x) {
        const int32_t r = rowidx[k];
        if (r < 0 || (int64_t)r >= nrow) atomicOr(bad, 2);
        else if (k > b && rowidx[k - 1] >= r) atomicOr(bad, 4);
    }
}

namespace svb {
void csc_validate(const svb_matrix_s *a, const char *who) {
    cudaStream_t st = ctx().stream;
    if (a->ncol <= 0) return;
    DevBuf<int> d_bad(1);
    SVB_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
    csc_validate_kernel<<<(unsigned)a->ncol, 128, 0, st>>>(a->colptr, a->rowidx, a->nrow, a->ncol, a->nnz, d_bad.p);
    count_launch();
    int bad = 0;
    SVB_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_CHECK(!(bad & 1), SVB_EDIM, std::string(who) + ": malformed colptr (not non-decreasing / out of range)");
    SVB_CHECK(!(bad & 2), SVB_EDIM, std::string(who) + ": a row index lies outside [1, nrow]");
    SVB_CHECK(!(bad & 4), SVB_EDIM, std::string(who) + ": row indices must ascend strictly inside every column (SparseMatrixCSC)");
}
}  // namespace svb

extern "C" {

int svb_csc_upload(int64_t nrow, int64_t ncol, const int64_t *colptr, const void *rowval, int rowval_type,
                   const void *nzval, int vtype, int index_base, svb_matrix_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(out != nullptr && colptr != nullptr, SVB_EARG, "svb_csc_upload: null pointer");
    SVB_CHECK(nrow >= 0 && ncol >= 0 && nrow < 2147483647LL, SVB_EDIM, "svb_csc_upload: bad dimensions (nrow must fit int32)");
    SVB_CHECK(index_base == 0 || index_base == 1, SVB_EARG, "index_base must be 0 or 1");
    SVB_CHECK(rowval_type == SVB_I32 || rowval_type == SVB_I64, SVB_EARG, "rowval_type must be I32 or I64");
    const int64_t nnz = colptr[ncol] - index_base;
    SVB_CHECK(nnz >= 0 && colptr[0] == index_base, SVB_EDIM, "svb_csc_upload: malformed colptr");
    SVB_CHECK(nnz == 0 || (rowval && nzval), SVB_EARG, "svb_csc_upload: null rowval/nzval");
    const int dev_vtype = (vtype == SVB_I64) ? SVB_I32 : vtype;
    cudaStream_t st = ctx().stream;
    svb_matrix_s *a = matrix_alloc(nrow, ncol, nnz, dev_vtype);
    try {
        DevBuf<int> d_over(1);
        SVB_CUDA(cudaMemsetAsync(d_over.p, 0, sizeof(int), st));
        // colptr
        SVB_CUDA(cudaMemcpyAsync(a->colptr, colptr, (size_t)(ncol + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        if (index_base)
            convert_offset_kernel<int64_t, int64_t><<<grid_for(ncol + 1), 256, 0, st>>>(a->colptr, a->colptr, ncol + 1, -index_base, nullptr);
        count_launch();
        if (nnz > 0) {
            if (rowval_type == SVB_I64)
                upload_index<int64_t>((const int64_t *)rowval, nnz, index_base, a->rowidx, d_over.p);
            else if (index_base == 0)
                SVB_CUDA(cudaMemcpyAsync(a->rowidx, rowval, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
            else
                upload_index<int32_t>((const int32_t *)rowval, nnz, index_base, a->rowidx, d_over.p);
            if (vtype == SVB_I64)
                upload_index<int64_t>((const int64_t *)nzval, nnz, 0, (int32_t *)a->val, d_over.p);
            else
                SVB_CUDA(cudaMemcpyAsync(a->val, nzval, (size_t)nnz * vtype_size(vtype), cudaMemcpyHostToDevice, st));
        }
        int over = 0;
        SVB_CUDA(cudaMemcpyAsync(&over, d_over.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        SVB_CHECK(!over, SVB_EDIM, "svb_csc_upload: an index or count does not fit in int32");
        svb::csc_validate(a, "svb_csc_upload");
    } catch (...) {
        delete a;
        throw;
    }
    *out = a;
    SVB_API_END
}

int svb_matrix_free(svb_matrix_t a) {
    SVB_API_BEGIN
    if (a) {
        if (ctx().initialised) cudaStreamSynchronize(ctx().stream);
        delete a;
    }
    SVB_API_END
}

int svb_matrix_info(svb_matrix_t a, int64_t *nrow, int64_t *ncol, int64_t *nnz, int *vtype) {
    SVB_API_BEGIN
    SVB_CHECK(a, SVB_EARG, "null matrix handle");
    if (nrow) *nrow = a->nrow;
    if (ncol) *ncol = a->ncol;
    if (nnz) *nnz = a->nnz;
    if (vtype) *vtype = a->vtype;
    SVB_API_END
}

int svb_matrix_download(svb_matrix_t a, int64_t *colptr, int64_t *rowval, void *nzval, int vtype, int index_base) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a, SVB_EARG, "null matrix handle");
    cudaStream_t st = ctx().stream;
    if (colptr) {
        SVB_CUDA(cudaMemcpyAsync(colptr, a->colptr, (size_t)(a->ncol + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        if (index_base)
            for (int64_t i = 0; i <= a->ncol; ++i) colptr[i] += index_base;
    }
    const int64_t chunk = std::min<int64_t>(std::max<int64_t>(a->nnz, 1), STAGE_ELEMS);
    if (rowval && a->nnz > 0) {
        DevBuf<int64_t> stage((size_t)chunk);
        for (int64_t o = 0; o < a->nnz; o += chunk) {
            const int64_t c = std::min(chunk, a->nnz - o);
            convert_offset_kernel<int32_t, int64_t><<<grid_for(c), 256, 0, st>>>(a->rowidx + o, stage.p, c, index_base, nullptr);
            count_launch();
            SVB_CUDA(cudaMemcpyAsync(rowval + o, stage.p, (size_t)c * 8, cudaMemcpyDeviceToHost, st));
        }
        SVB_CUDA(cudaStreamSynchronize(st));
    }
    if (nzval && a->nnz > 0) {
        if (vtype == a->vtype) {
            SVB_CUDA(cudaMemcpyAsync(nzval, a->val, (size_t)a->nnz * vtype_size(vtype), cudaMemcpyDeviceToHost, st));
        } else {
            DevBuf<char> stage((size_t)chunk * vtype_size(vtype));
            for (int64_t o = 0; o < a->nnz; o += chunk) {
                const int64_t c = std::min(chunk, a->nnz - o);
                const unsigned g = grid_for(c);
                if (a->vtype == SVB_I32 && vtype == SVB_I64)
                    convert_value_kernel<int32_t, int64_t><<<g, 256, 0, st>>>((const int32_t *)a->val + o, (int64_t *)stage.p, c);
                else if (a->vtype == SVB_I32 && vtype == SVB_F64)
                    convert_value_kernel<int32_t, double><<<g, 256, 0, st>>>((const int32_t *)a->val + o, (double *)stage.p, c);
                else if (a->vtype == SVB_F32 && vtype == SVB_F64)
                    convert_value_kernel<float, double><<<g, 256, 0, st>>>((const float *)a->val + o, (double *)stage.p, c);
                else if (a->vtype == SVB_F64 && vtype == SVB_F32)
                    convert_value_kernel<double, float><<<g, 256, 0, st>>>((const double *)a->val + o, (float *)stage.p, c);
                else
                    throw Error(SVB_EARG, "svb_matrix_download: unsupported value conversion");
                count_launch();
                SVB_CUDA(cudaMemcpyAsync((char *)nzval + (size_t)o * vtype_size(vtype), stage.p, (size_t)c * vtype_size(vtype),
                                         cudaMemcpyDeviceToHost, st));
            }
        }
        SVB_CUDA(cudaStreamSynchronize(st));
    }
    SVB_API_END
}

int svb_column_subset(svb_matrix_t a, const int64_t *idx, int64_t k, int index_base, svb_matrix_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && out && (idx || k == 0), SVB_EARG, "svb_column_subset: null argument");
    SVB_CHECK(k >= 0, SVB_EDIM, "svb_column_subset: negative k");
    std::vector<int64_t> h((size_t)k);
    for (int64_t i = 0; i < k; ++i) {
        h[i] = idx[i] - index_base;
        SVB_CHECK(h[i] >= 0 && h[i] < a->ncol, SVB_EDIM, "svb_column_subset: column index out of range");
    }
    cudaStream_t st = ctx().stream;
    DevBuf<int64_t> d_idx((size_t)std::max<int64_t>(k, 1));
    DevBuf<int64_t> lens((size_t)(k + 1));
    if (k) SVB_CUDA(cudaMemcpyAsync(d_idx.p, h.data(), (size_t)k * 8, cudaMemcpyHostToDevice, st));
    subset_len_kernel<<<(unsigned)((k + 256) / 256), 256, 0, st>>>(a->colptr, d_idx.p, k, lens.p);
    count_launch();
    SVB_LAUNCH_CHECK();
    *out = finish_subset(a, lens, k, a->nrow, d_idx.p, 0, nullptr);
    SVB_API_END
}

int svb_row_slice(svb_matrix_t a, int64_t row0, int64_t row1, svb_matrix_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && out, SVB_EARG, "svb_row_slice: null argument");
    SVB_CHECK(0 <= row0 && row0 <= row1 && row1 <= a->nrow, SVB_EDIM, "svb_row_slice: bad row range");
    cudaStream_t st = ctx().stream;
    DevBuf<int64_t> srcbeg((size_t)std::max<int64_t>(a->ncol, 1));
    DevBuf<int64_t> lens((size_t)(a->ncol + 1));
    slice_bounds_kernel<<<(unsigned)((a->ncol + 256) / 256), 256, 0, st>>>(a->colptr, a->rowidx, a->ncol, row0, row1, srcbeg.p, lens.p);
    count_launch();
    SVB_LAUNCH_CHECK();
    *out = finish_subset(a, lens, a->ncol, row1 - row0, nullptr, (int32_t)row0, srcbeg.p);
    SVB_API_END
}

int svb_transpose(svb_matrix_t a, svb_matrix_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(a && out, SVB_EARG, "svb_transpose: null argument");
    SVB_CHECK(a->ncol < 2147483647LL, SVB_EDIM, "svb_transpose: ncol must fit int32");
    *out = matrix_transpose(a);
    SVB_API_END
}

}  // extern "C"

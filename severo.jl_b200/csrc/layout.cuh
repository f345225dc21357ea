// layout.cuh — device layouts of the operator and their builders (matrix.cu).
#pragma once
#include "svb_internal.h"

namespace svb {

// cells tiled by R = 2^log2R; inside a tile the nonzeros are gene-major (ascending gene, then cell)
template <typename V>
struct TileCSC {
    int64_t R = 0, ntiles = 0;
    int log2R = 0;
    DevBuf<int64_t> gptr;   // [ntiles*ncol + 1]
    DevBuf<uint16_t> rloc;  // [nnz] cell index inside the tile
    DevBuf<V> aval;         // [nnz]
};

template <typename VI, typename VO>
void build_tilecsc(const svb_matrix_s *a, int log2R, TileCSC<VO> &out);
template <typename VI, typename VO>
void build_tilecsc_from_transposed(const svb_matrix_s *a, int log2R, TileCSC<VO> &out);
template <typename V, typename IdxT>
void csr_from_tilecsc(const TileCSC<V> &tc, const svb_matrix_s *a, DevBuf<int64_t> &rowptr, DevBuf<IdxT> &fidx,
                      DevBuf<V> &fval);
// startpos[t*ncol + j] = first position of column j whose row >= t*2^log2R, t = 0..ntiles (inclusive)
void tile_bounds(const svb_matrix_s *a, int64_t tile_rows, int64_t ntiles, int64_t *startpos);
svb_matrix_s *matrix_transpose(const svb_matrix_s *a);
void launch_strided_copy(const int64_t *src, int64_t stride, int64_t n, int64_t *dst, cudaStream_t st);

}  // namespace svb

// spmul.cu — the stand-alone sparse products of src/mul.jl:
//   mul!(y, A::CSC, x::SparseVector, alpha, beta)      mul.jl:50-77   y = beta*y + alpha*A*x        (SpMSpV)
//   mul!(C::Dense, A::CSC, B::CSC, alpha, beta)        mul.jl:82-114  C = beta*C + alpha*A*B        (CSC x CSC -> dense)
//   the Transpose / Adjoint methods                    mul.jl:79-80   mul!(C, copy(A'), B, alpha, beta)
// The reference scatters: for every stored x_j (ascending j) it walks column j of A and does y[i] += (alpha*a_ij)*x_j — each
// y[i] therefore receives its terms in ASCENDING j, every term rounded as (alpha*a)*x and added one at a time. A parallel
// scatter would need floating-point atomics (another order on every run); here the product is turned round: y[i] is one
// thread's sequential sum over the stored entries of ROW i of A (ascending j, the layout of the device transpose A' — or A
// itself for the Adjoint methods, which need the rows of A'), taking only the j that x stores. Same terms, same order, same
// roundings: the results are bit-identical to the reference loop and run-to-run deterministic. Stored zeros of x take part,
// as in the reference (a mask says "stored", not the value).
#include "svb_internal.h"
#include "layout.cuh"

#include <algorithm>
#include <vector>

using namespace svb;

namespace svb {

__device__ __forceinline__ double val_as_f64(const void *val, int vtype, int64_t k) {
    if (vtype == SVB_F64) return ((const double *)val)[k];
    if (vtype == SVB_F32) return (double)((const float *)val)[k];
    return (double)((const int32_t *)val)[k];
}

// xd[j*ldx + c], stored[j*ldx + c] <- the stored entries of the sparse columns (dense scratch, one column per right-hand side)
__global__ void spm_densify_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx, const void *__restrict__ val,
                                   int vtype, int64_t col0, int ncols, int ldx, double *__restrict__ xd,
                                   uint8_t *__restrict__ stored) {
    const int c = blockIdx.y;
    if (c >= ncols) return;
    const int64_t b = ptr[col0 + c], e = ptr[col0 + c + 1];
    for (int64_t k = b + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < e; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = idx[k];
        xd[j * ldx + c] = val_as_f64(val, vtype, k);
        stored[j * ldx + c] = 1;
    }
}

// rows of A given as the CSC of A' (rptr / cidx ascending inside a row). One thread per (row, right-hand side).
// out[i + c*ldo] = beta-scaled old value (mul.jl:55-57 / :97-99) + sum over the row's entries whose column x stores.
__global__ void __launch_bounds__(256) spm_rows_kernel(const int64_t *__restrict__ rptr, const int32_t *__restrict__ cidx,
                                                       const void *__restrict__ aval, int vtype, int64_t m, int ncols, int ldx,
                                                       const double *__restrict__ xd, const uint8_t *__restrict__ stored,
                                                       double alpha, double beta, int alpha0_returns, double *__restrict__ out, int64_t ldo) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / ncols;
    const int c = (int)(t - i * ncols);
    if (i >= m) return;
    double acc = out[i + (int64_t)c * ldo];
    if (beta != 1.0) acc = (beta == 0.0) ? 0.0 : __dmul_rn(acc, beta);  // fill!(y, 0) / rmul!(y, beta)
    if (!(alpha0_returns && alpha == 0.0)) {                             // mul.jl:58 "alpha == 0 && return y" (SpMSpV only)
        for (int64_t k = rptr[i]; k < rptr[i + 1]; ++k) {
            const int64_t j = cidx[k];
            if (stored[j * ldx + c]) {
                const double term = __dmul_rn(__dmul_rn(alpha, val_as_f64(aval, vtype, k)), xd[j * ldx + c]);  // (alpha*a)*x
                acc = __dadd_rn(acc, term);
            }
        }
    }
    out[i + (int64_t)c * ldo] = acc;
}

// C (m x p) += rows(A) x sparse columns [col0, col0 + ncols) of (bptr, bidx, bval); rows as the CSC of A'
static void rows_times_sparse_cols(const svb_matrix_s *rowsA, int64_t m, int64_t n, const int64_t *bptr, const int32_t *bidx,
                                   const void *bval, int bvtype, int64_t p, double alpha, double beta, int alpha0_returns, double *dC,
                                   int64_t ldc) {
    cudaStream_t st = ctx().stream;
    constexpr int KC = 8;  // right-hand sides per pass over the rows
    DevBuf<double> xd((size_t)std::max<int64_t>(n, 1) * KC);
    DevBuf<uint8_t> stored((size_t)std::max<int64_t>(n, 1) * KC);
    for (int64_t c0 = 0; c0 < p; c0 += KC) {
        const int kc = (int)std::min<int64_t>(KC, p - c0);
        SVB_CUDA(cudaMemsetAsync(stored.p, 0, (size_t)std::max<int64_t>(n, 1) * KC, st));
        dim3 grid(64, (unsigned)kc);
        spm_densify_kernel<<<grid, 256, 0, st>>>(bptr, bidx, bval, bvtype, c0, kc, KC, xd.p, stored.p);
        const int64_t threads = m * kc;
        spm_rows_kernel<<<(unsigned)std::max<int64_t>(1, (threads + 255) / 256), 256, 0, st>>>(
            rowsA->colptr, rowsA->rowidx, rowsA->val, rowsA->vtype, m, kc, KC, xd.p, stored.p, alpha, beta, alpha0_returns, dC + c0 * ldc, ldc);
        count_launch(2);
        SVB_LAUNCH_CHECK();
    }
}

}  // namespace svb

extern "C" {

int svb_spmspv(svb_matrix_t A, const int64_t *x_nzind, const double *x_nzval, int64_t nx, int index_base, double alpha, double beta,
               double *y) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(A && y && (nx == 0 || (x_nzind && x_nzval)), SVB_EARG, "svb_spmspv: null argument");
    SVB_CHECK(nx >= 0 && nx <= A->ncol, SVB_EDIM, "svb_spmspv: x stores more entries than its length (DimensionMismatch)");
    const int64_t m = A->nrow, n = A->ncol;
    if (m == 0) return SVB_OK;
    std::vector<int32_t> idx((size_t)std::max<int64_t>(nx, 1));
    for (int64_t k = 0; k < nx; ++k) {
        const int64_t j = x_nzind[k] - index_base;
        SVB_CHECK(j >= 0 && j < n, SVB_EDIM, "svb_spmspv: index of x out of range (DimensionMismatch)");
        SVB_CHECK(k == 0 || j > idx[(size_t)k - 1], SVB_EDIM, "svb_spmspv: the indices of x must ascend strictly (SparseVector)");
        idx[(size_t)k] = (int32_t)j;
    }
    cudaStream_t st = ctx().stream;
    struct Owned {
        svb_matrix_s *p = nullptr;
        ~Owned() { delete p; }
    } rows;
    rows.p = matrix_transpose(A);  // rows of A, ascending column inside a row
    DevBuf<int64_t> xptr(2);
    DevBuf<int32_t> xidx((size_t)std::max<int64_t>(nx, 1));
    DevBuf<double> xval((size_t)std::max<int64_t>(nx, 1)), dy((size_t)m);
    const int64_t hp[2] = {0, nx};
    SVB_CUDA(cudaMemcpyAsync(xptr.p, hp, sizeof(hp), cudaMemcpyHostToDevice, st));
    if (nx) {
        SVB_CUDA(cudaMemcpyAsync(xidx.p, idx.data(), (size_t)nx * 4, cudaMemcpyHostToDevice, st));
        SVB_CUDA(cudaMemcpyAsync(xval.p, x_nzval, (size_t)nx * 8, cudaMemcpyHostToDevice, st));
    }
    SVB_CUDA(cudaMemcpyAsync(dy.p, y, (size_t)m * 8, cudaMemcpyHostToDevice, st));
    rows_times_sparse_cols(rows.p, m, n, xptr.p, xidx.p, xval.p, SVB_F64, 1, alpha, beta, 1, dy.p, m);
    SVB_CUDA(cudaMemcpyAsync(y, dy.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_API_END
}

int svb_spgemm_dense(svb_matrix_t A, int transpose_a, svb_matrix_t B, double alpha, double beta, double *C, int64_t ldc) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(A && B && C, SVB_EARG, "svb_spgemm_dense: null argument");
    const int64_t m = transpose_a ? A->ncol : A->nrow, inner = transpose_a ? A->nrow : A->ncol, p = B->ncol;
    SVB_CHECK(inner == B->nrow, SVB_EDIM, "svb_spgemm_dense: size(A, 2) != size(B, 1) (DimensionMismatch)");
    SVB_CHECK(ldc >= std::max<int64_t>(m, 1), SVB_EDIM, "svb_spgemm_dense: leading dimension of C smaller than size(A, 1)");
    if (m == 0 || p == 0) return SVB_OK;
    cudaStream_t st = ctx().stream;
    struct Owned {
        svb_matrix_s *p = nullptr;
        ~Owned() { delete p; }
    } tr;
    const svb_matrix_s *rows = A;  // Adjoint / Transpose methods (mul.jl:79-80): the rows of A' are the columns of A
    if (!transpose_a) {
        tr.p = matrix_transpose(A);
        rows = tr.p;
    }
    DevBuf<double> dC((size_t)ldc * p);
    SVB_CUDA(cudaMemcpyAsync(dC.p, C, (size_t)ldc * p * 8, cudaMemcpyHostToDevice, st));
    rows_times_sparse_cols(rows, m, inner, B->colptr, B->rowidx, B->val, B->vtype, p, alpha, beta, 0, dC.p, ldc);
    SVB_CUDA(cudaMemcpyAsync(C, dC.p, (size_t)ldc * p * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_API_END
}

}  // extern "C"

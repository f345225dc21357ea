/*
 * severo_b200.h — C ABI of libsevero_b200.so: hand-written sm_100a CUDA kernels for the
 * truncated-PCA hot path of ExaScience/Severo.jl (log-normalise -> per-gene moments ->
 * centre/scale -> IRLBA on the implicit centred sparse operator).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ / torch types. It replaces
 *   - the `ccall(("irlba", libcell), ...)` of src/irlba.jl:66-71 together with the two Julia
 *     callbacks `matmul` (src/irlba.jl:23-38) and `randv` (src/irlba.jl:40-45): the mat-vec
 *     now runs on the GPU inside svb_irlba, nothing calls back into the host language;
 *   - the Julia loops of src/normalize.jl:17-38, src/scaling.jl:18-34,119-147,199-217,245-272
 *     and src/variablefeatures.jl:19-28 (one entry point each, cited below).
 * The reference-side bindings (Julia `ccall` overlay, Python ctypes) are shown in INTEGRATION.md.
 *
 * Conventions
 *   - svb_init(device): the calling process drives ONE GPU; N GPUs = N processes, each holding a
 *     contiguous range of cells, joined by svb_comm_init (NCCL over NVLink). OR svb_init_devices:
 *     one process and one calling thread drive N GPUs through worker threads inside the library.
 *   - Matrices follow Severo's orientation: rows = cells, columns = genes, CSC (column = gene),
 *     exactly the layout of `SparseMatrixCSC{T,Int64}` (src/Severo.jl:26-34). Index arrays may be
 *     Julia's 1-based Int64 (index_base = 1) or 0-based.
 *   - Dense arrays are column-major Float64 (Julia `Matrix{Float64}`).
 *   - Every function returns 0 on success or a negative code; svb_last_error() gives the text.
 *     Codes -1/-2/-3/-4 keep the meaning of the legacy `irlba` return value (src/irlba.jl:73
 *     treats any non-zero as "convergence failed").
 *   - Calls are synchronous for the caller; the library owns its CUDA stream (or uses the one
 *     given to svb_set_stream). Not re-entrant; call from one host thread at a time.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SEVERO_B200_H
#define SEVERO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct svb_matrix_s *svb_matrix_t;     /* device-resident sparse matrix, CSC cells x genes */
typedef struct svb_operator_s *svb_operator_t; /* implicit operator S = A - 1*mu' (scaling.jl:219-232) */
typedef struct svb_result_s *svb_result_t;     /* device-resident U, s, V of one IRLBA solve */

/* element types of host/device value arrays */
#define SVB_I32 0
#define SVB_I64 1
#define SVB_F32 2
#define SVB_F64 3

/* return codes */
#define SVB_OK 0
#define SVB_EDIM (-1)       /* bad dimensions / arguments */
#define SVB_ENOCONV (-2)    /* not converged within maxit */
#define SVB_ENOMEM (-3)     /* out of (device) memory */
#define SVB_ENULLSPACE (-4) /* starting vector (numerically) in the null space */
#define SVB_EARG (-5)       /* invalid handle / enum / null pointer */
#define SVB_ECUDA (-100)    /* CUDA runtime failure (also: no device) */
#define SVB_ENCCL (-101)    /* NCCL failure / NCCL not loadable */

/* normalisation methods (normalize.jl:41-50) */
#define SVB_NORM_LOGNORMALIZE 0
#define SVB_NORM_RELATIVECOUNTS 1

/* ---- library / device -------------------------------------------------------------------- */
int svb_init(int device);
int svb_shutdown(void);
const char *svb_last_error(void);
const char *svb_version(void);
int svb_set_stream(void *cuda_stream); /* NULL = library-owned stream */
int svb_synchronize(void);
int svb_device_info(int *sm_count, int64_t *total_mem, int *cc_major, int *cc_minor);

/* per-kernel-class device timers (CUDA events on the launching stream). class ids below. */
#define SVB_K_SPMV_FWD 0   /* S*v   (scaling.jl:245-250) */
#define SVB_K_SPMV_ADJ 1   /* S'*w  (scaling.jl:252-257), incl. the partial reduce */
#define SVB_K_REORTH 2     /* tall-skinny CGS passes (libcell `orthog`) */
#define SVB_K_RESTART 3    /* W*P / V*Q restart + final GEMMs */
#define SVB_K_VECTOR 4     /* norms / scales / small vector kernels */
#define SVB_K_COMM 5       /* NCCL allreduce */
#define SVB_K_NCLASS 6
int svb_profile_enable(int on);
int svb_profile_reset(void);
/* ms[c], launches[c], bytes[c] (algorithmic bytes, SURVEY 8d formulas with this build's storage widths) */
int svb_profile_get(double *ms, int64_t *launches, double *bytes);
int64_t svb_launch_count(void); /* kernels launched by this library since init / last reset */
int svb_launch_count_reset(void);

/* ---- multi-GPU: one process per GPU, cells sharded, NCCL allreduce -------------------------- */
int svb_comm_unique_id(unsigned char id[128]);
int svb_comm_init(int nranks, int rank, const unsigned char id[128]);
int svb_comm_destroy(void);
int svb_comm_info(int *nranks, int *rank);
int svb_comm_allreduce_f64(double *host_buf, int64_t n); /* sum over ranks, in place (host buffer) */

/* ---- multi-GPU: ONE process, ONE calling thread, N GPUs (csrc/multi.cu) ---------------------------------------------------- */
/* The reference's boundary is one synchronous call from one host thread (src/irlba.jl:66-71); these entry points keep that
 * shape on a whole node: svb_init_devices starts one worker thread per GPU inside the library (devices = NULL: 0..ndev-1;
 * ndev <= 0: every visible GPU), joins them with ncclCommInitAll and maps their mailboxes through peer access. The
 * *_devices calls then take the caller's WHOLE SparseMatrixCSC on the host, shard it by cells internally, solve on all the
 * GPUs and fill the caller's s[nu], U[m x nu], V[n x nu] (column-major) — the Julia drop-in reaches 8 GPUs without a launcher.
 * Independent of svb_init (the one-GPU context of the calling thread); both may be active. */
int svb_init_devices(int ndev, const int *devices);
int svb_devices_info(int *ndev, int *devices, int *peer_mailboxes);
int svb_shutdown_devices(void);
/* irlba(CenteredMatrix(A, mu), nu) for a host SparseMatrixCSC{Float64|Float32} A (m cells x n genes; rowval Int32 or Int64,
 * index_base 1 for Julia), mu[n] or NULL: same buffers and return convention as svb_irlba. */
int svb_irlba_csc_devices(int64_t m, int64_t n, const int64_t *colptr, const void *rowval, int rowval_type,
                          const void *nzval, int vtype, int index_base, const double *mu, int64_t nu,
                          int64_t m_b, int64_t maxit, double tol, double svtol, const double *init,
                          double *s, double *U, double *V, int64_t *iter, int64_t *mprod);
/* The fused PCA call over the RAW COUNTS of the HVG columns (svb_operator_create_counts on every device, moments over the
 * cells of all devices): counts Int32 / Int64 CSC (m x n), libsize[m] = library sizes of the full matrix; mu_out[n]
 * (optional) receives the stored centre mean/sd. */
int svb_pca_counts_devices(int64_t m, int64_t n, const int64_t *colptr, const void *rowval, int rowval_type,
                           const void *counts, int vtype, int index_base, const int64_t *libsize,
                           double scale_factor, double scale_max, int64_t nu, int64_t m_b, int64_t maxit,
                           double tol, double svtol, const double *init, double *mu_out, double *s,
                           double *U, double *V, int64_t *iter, int64_t *mprod);

/* ---- sparse matrices (CSC, cells x genes) ----------------------------------------------------- */
/* Upload a SparseMatrixCSC. colptr: int64[ncol+1]. rowval: int32 or int64 (rowval_type), nzval:
 * vtype in {I32,I64,F32,F64}. index_base 1 for Julia arrays. Int64 values are narrowed to int32 on
 * the device (counts; overflow is an error), Float32 stays Float32, Float64 stays Float64. */
int svb_csc_upload(int64_t nrow, int64_t ncol, const int64_t *colptr, const void *rowval,
                   int rowval_type, const void *nzval, int vtype, int index_base,
                   svb_matrix_t *out);
/* The e2e form of "upload the HVG counts, then scale_features' moments" (normalize.jl:27,36 + scaling.jl:18-34 on Y[:, hvf]): the
 * SparseMatrixCSC{Int32,Int64|Int32} of the HVG counts is uploaded column group by column group, DENSEST GENES FIRST, and the
 * order-exact Welford chains of the groups that have arrived (a serial chain per gene: ~0.14 us per stored value, 0.15 s for the
 * densest gene of a 1.3 M-cell matrix) run on side streams while the rest of the matrix is still crossing PCIe. mean / var [ncol]
 * are bit-identical to svb_normalize_libsize + svb_mean_var; out = the device matrix of the counts (as svb_csc_upload). */
int svb_csc_upload_lognorm_moments(int64_t nrow, int64_t ncol, const int64_t *colptr, const void *rowval,
                                   int rowval_type, const int32_t *counts, int index_base,
                                   const int64_t *libsize, double scale_factor, double *mean, double *var,
                                   svb_matrix_t *out);
int svb_matrix_free(svb_matrix_t a);
int svb_matrix_info(svb_matrix_t a, int64_t *nrow, int64_t *ncol, int64_t *nnz, int *vtype);
/* Download. Any of colptr/rowval/nzval may be NULL. Indices are written as int64 with index_base. */
int svb_matrix_download(svb_matrix_t a, int64_t *colptr, int64_t *rowval, void *nzval, int vtype,
                        int index_base);
/* X[:, idx] (docs/src/pbmc.md:121; scaling.jl:339). idx in any order, duplicates allowed. */
int svb_column_subset(svb_matrix_t a, const int64_t *idx, int64_t k, int index_base,
                      svb_matrix_t *out);
/* rows [row0, row1) of a (0-based, half-open): the cell shard of one rank. */
int svb_row_slice(svb_matrix_t a, int64_t row0, int64_t row1, svb_matrix_t *out);
/* copy(X') : stable transpose on the device (every reader of src/input.jl ends with it). */
int svb_transpose(svb_matrix_t a, svb_matrix_t *out);

/* filtering.jl:15-35,101-106 filter_cells, then filter_features on the remaining cells (= filter_counts; with
 * min_cells = 0 it is filter_cells alone, with the three cell thresholds 0 it is filter_features alone):
 *   cell kept    <=> #{features with count > min_feature_count} >= min_features  and (min_umi <= 0 or UMI total > min_umi)
 *   feature kept <=> #{kept cells with count > 0} >= min_cells
 * cell_keep[nrow] / feature_keep[ncol] (optional) receive the 0/1 masks (the reference's CI / FI); out = A[CI, FI] with
 * every stored entry of a kept pair, rows renumbered in order. Integer counts only. */
int svb_filter_counts(svb_matrix_t counts, int64_t min_cells, int64_t min_features,
                      int64_t min_feature_count, int64_t min_umi, uint8_t *cell_keep,
                      uint8_t *feature_keep, svb_matrix_t *out);

/* ---- pre-processing sweeps ---------------------------------------------------------------- */
/* normalize.jl:17-55  row_norm / log_norm. Integer counts in; dtype = SVB_F32 | SVB_F64 out.
 * B = scale_factor * x / s_cell (one rounded multiply, one rounded divide), then log1p. */
int svb_normalize(svb_matrix_t counts, int method, double scale_factor, int dtype,
                  svb_matrix_t *out);
/* The same sweep on a column subset of the counts (the HVG columns): the library sizes of the FULL matrix are
 * supplied (libsize[nrow], from svb_row_sums) instead of being recomputed, so the values equal Y[:, hvf]. */
int svb_normalize_libsize(svb_matrix_t counts_subset, const int64_t *libsize, int method,
                          double scale_factor, int dtype, svb_matrix_t *out);
/* normalize.jl:24 s = sum(A, dims=2): exact int64 library sizes (length nrow). With a communicator
 * nothing is exchanged: a cell lives on one rank. */
int svb_row_sums(svb_matrix_t counts, int64_t *s);
/* scaling.jl:18-34,132-142 mean_var(A): order-exact sequential Welford per gene. */
int svb_mean_var(svb_matrix_t a, double *mu, double *var);
/* The same Welford chain continued across cell shards: count/mu/s (length ncol) hold the state after all
 * cells of the previous ranks (first rank: count = total cells - total nonzeros of the gene, mu = s = 0,
 * scaling.jl:21) and receive the state after this shard; var = s/(cells-1) after the last rank. Bit-identical
 * to one sequential pass over the unsharded matrix. */
int svb_welford_carry(svb_matrix_t a, int64_t *count, double *mu, double *s);
/* variablefeatures.jl:19-28 standardized_var_clipped (vmax <= 0 => sqrt(nrow)). */
int svb_stdvar_clipped(svb_matrix_t counts, const double *mu, const double *sd, double vmax,
                       double *out);
/* scaling.jl:199-217 scale_data: out = min(x/std, scale_max + mu/std); mu_out = mean/std. */
int svb_scale(svb_matrix_t a, double scale_max, int dtype, svb_matrix_t *out, double *mu_out);
/* Same sweep with caller-supplied per-gene mean / unbiased variance: the cell-sharded form, where the
 * moments of the whole matrix come from merging the per-rank moments (SURVEY 8e "fast mode"). */
int svb_scale_with_moments(svb_matrix_t a, const double *mean, const double *var, double scale_max,
                           int dtype, svb_matrix_t *out, double *mu_out);

/* ---- the implicit centred operator (scaling.jl:219-272) -------------------------------------- */
/* S = A - 1*mu' with A = a (transposed = 0) or A = a' (transposed = 1, the lazy Adjoint of
 * test_irlba.jl:111). mu may be NULL (plain sparse matrix). Builds the two streaming layouts
 * (row-major for S*v, cell-tiled gene-major for S'*w) on the device; `a` is not retained. */
int svb_operator_create(svb_matrix_t a, const double *mu, int transposed, svb_operator_t *out);
/* Same, choosing the storage width of the values in the two streaming layouts: 0 = as the input,
 * SVB_F32 = Float32 storage with Float64 accumulation (the optional 1e-4 mode: 6 instead of 10 bytes
 * per nonzero), SVB_F64. */
int svb_operator_create_ex(svb_matrix_t a, const double *mu, int transposed, int value_storage,
                           svb_operator_t *out);
/* Dense column-major A (m x n, leading dimension lda >= m) for the StridedMatrix methods. */
int svb_operator_create_dense(int64_t m, int64_t n, const double *a, int64_t lda, const double *mu,
                              int transposed, svb_operator_t *out);
/* The same operator built from the RAW COUNTS of the HVG columns — the scaled matrix is never materialised.
 * Replaces scale_features + CenteredMatrix + the normalised values (normalize.jl:27,36; scaling.jl:199-232) for the
 * PCA call: entry (i,j) of S is min(log1p(scale_factor*c_ij/libsize_i)/sd_j, scale_max + mean_j/sd_j) - mean_j/sd_j,
 * stored as ONE 16-bit code per nonzero (2 instead of 10 bytes); the value is rebuilt in the product kernels from a
 * per-cell table over the count levels 1..levels and the per-gene 1/sd. Counts above `levels`, counts < 1 and
 * clipped entries keep their exact Float64 value as exception chunks of the same streams (17 B per such entry). counts: int32 CSC, cells x HVGs (<= 65534
 * genes). libsize[nrow]: library sizes of the FULL count matrix (svb_row_sums). mean/var[ncol]: moments of the
 * log-normalised HVG columns (svb_mean_var), or both NULL: they are then computed here from the counts by two parallel
 * passes (with a communicator: over the cells of all ranks) — a few ulp from the sequential Welford of scaling.jl:18-34,
 * for the fused PCA call that keeps them internal. levels: 0 = choose (4, 8, 16 or 32). mu_out (optional, [ncol])
 * receives the stored centre mean/sd (scaling.jl:207). Entries are within 2 ulp of the reference's (t*(1/sd) vs t/sd). */
int svb_operator_create_counts(svb_matrix_t counts, const int64_t *libsize, double scale_factor,
                               const double *mean, const double *var, double scale_max, int levels,
                               double *mu_out, svb_operator_t *out);
/* levels L, cells per adjoint tile, entries stored as codes / as exception chunks, 16-byte chunks of the two streams */
int svb_operator_counts_info(svb_operator_t op, int *levels, int64_t *tile_cells, int64_t *nnz_coded,
                             int64_t *nnz_exception, int64_t *fwd_chunks, int64_t *adj_chunks);
/* Shared-memory layout of the count-level operator's gathered tables: the forward kernel holds fwd_replicas bank-shifted copies
 * of x/sd, the adjoint kernel adj_replicas copies of the first adj_replicated_levels levels of its tile table; *_passes = the
 * average number of shared-memory passes per set of 16 gathers after the build-time replica assignment (1 = conflict-free;
 * 0 = not evaluated: no replicas and SVB_FACT_STATS unset). Any pointer may be NULL. */
int svb_operator_counts_layout(svb_operator_t op, int *fwd_replicas, double *fwd_passes,
                               int *adj_replicas, int *adj_replicated_levels, double *adj_passes);
/* Debug / study aid: copies one stream of the count-level operator to the host — adjoint = 0: the forward (cell-major) stream,
 * 1: the adjoint (tile, gene) stream. code[chunks*8] (16-bit codes, or the 16 bytes of an exception chunk), meta[chunks].
 * Either pointer may be NULL. Used by the layout tests and tools/studies/. */
int svb_operator_counts_stream(svb_operator_t op, int adjoint, uint16_t *code, uint8_t *meta);
int svb_operator_free(svb_operator_t op);
int svb_operator_info(svb_operator_t op, int64_t *m, int64_t *n, int64_t *nnz, int *is_dense,
                      int *value_bytes, int *index_bytes);
/* mul!(C, S, v, alpha, beta) / mul!(C, S', v, alpha, beta), vector and k-column matrix forms
 * (scaling.jl:245-272). trans = 'N' or 'T' (84, as the legacy matmul callback). Host buffers,
 * column-major with leading dimension = vector length. The adjoint matrix form subtracts the
 * rank-1 term (the reference's :271 adds it — untested sign slip). */
int svb_mul(svb_operator_t op, char trans, double alpha, const double *x, double beta, double *y,
            int64_t k);
/* Device-pointer variant for benchmarking one product (x, y already in HBM). */
int svb_mul_device(svb_operator_t op, char trans, double alpha, const double *dx, double beta,
                   double *dy);

/* ---- the stand-alone sparse products of src/mul.jl ------------------------------------------------------------- */
/* mul.jl:50-77  mul!(y, A::SparseMatrixCSC, x::SparseVector, alpha, beta): y = beta*y + alpha*A*x. A: m x n handle (values of any
 * uploaded type, promoted to Float64); x: its nx stored entries (x_nzind ascending, index_base 1 for Julia's nonzeroinds; stored
 * zeros take part, as in the reference); y: host Float64[m], in and out. Same terms, same order, same roundings as the reference
 * loop ((alpha*a)*x added to y[i] in ascending j): bit-identical and deterministic (csrc/spmul.cu). */
int svb_spmspv(svb_matrix_t A, const int64_t *x_nzind, const double *x_nzval, int64_t nx, int index_base,
               double alpha, double beta, double *y);
/* mul.jl:82-114  mul!(C::StridedMatrix, A::CSC, B::CSC, alpha, beta): C = beta*C + alpha*A*B, C host Float64 column-major
 * (size(A,1) x size(B,2), leading dimension ldc). transpose_a != 0: the Transpose / Adjoint methods of mul.jl:79-80,
 * C = beta*C + alpha*A'*B (the reference materialises copy(A'); here the rows of A' are the columns of A as stored).
 * Bit-identical to the reference's triple loop (every C[row, col] receives (alpha*a)*b in ascending inner index). */
int svb_spgemm_dense(svb_matrix_t A, int transpose_a, svb_matrix_t B, double alpha, double beta,
                     double *C, int64_t ldc);

/* C'C (scaling.jl:274-296, over the CSC x CSC -> dense product of mul.jl:82-114): G (n x n, column-major, host) =
 * S'S = A'A - mu q' - q mu' + M mu mu', q = column sums of A, M = cells. Explicit sparse operators: one fused pass per
 * block of 4 genes over the adjoint layout (lower triangle only, fixed summation order, exactly symmetric result);
 * dense and count-level operators: column j = S'(S e_j). With a communicator G is summed over the cell shards and
 * identical on all ranks. */
int svb_gram(svb_operator_t op, double *G);

/* ---- IRLBA (replaces libcell `irlba`, src/irlba.jl:66-71) -------------------------------------- */
/* Same buffers and return convention as the legacy call: init[n] in; s[nu], U[m x nu], V[n x nu]
 * out (column-major, caller-owned). m_b = work size (irlba.jl:50: nu+7, clamped to min(m,n)).
 * restart > 0: the first `restart` columns/values of U, s, V are inputs (irlba.jl:93-98).
 * With a communicator, m is the local number of cells and U the local rows; s and V are
 * identical on all ranks. iter / mprod (optional) return restarts and mat-vec count. */
int svb_irlba(svb_operator_t op, int64_t nu, int64_t m_b, int64_t maxit, int64_t restart,
              double tol, double svtol, const double *init, double *s, double *U, double *V,
              int64_t *iter, int64_t *mprod);
/* Split form: solve leaving U, s, V on the device; download later (or never, for timing). */
int svb_irlba_solve(svb_operator_t op, int64_t nu, int64_t m_b, int64_t maxit, int64_t restart,
                    double tol, double svtol, const double *init, const double *s0,
                    const double *U0, const double *V0, svb_result_t *out);
int svb_result_info(svb_result_t r, int64_t *m, int64_t *n, int64_t *nu, int64_t *iter,
                    int64_t *mprod, int *info);
/* scale_u != 0 returns Z = U*Diagonal(s) (embedding.jl:67 coordinates) instead of U. */
int svb_result_download(svb_result_t r, double *s, double *U, double *V, int scale_u);
int svb_result_free(svb_result_t r);

/* tssvd (embedding.jl:30-44, `embedding(...; algorithm=:tssvd)`): C = Hermitian(A'A) on the device (svb_gram's kernels),
 * its nsv largest eigenpairs (lambda, phi) by the device solver applied to C (the reference calls Arpack `eigs` on the
 * host), Sigma = sqrt(lambda), U = A*phi*inv(Diagonal(Sigma)). ncv = Lanczos basis size (the reference's default is
 * 2*nsv; <= 0 picks nsv+7), tol <= 0 (eigs' tol = 0.0, machine precision) is served at 1e-12, init[n] = start vector. Result as for
 * svb_irlba_solve (U local rows with a communicator); info = -2 when not converged within maxit. */
int svb_tssvd(svb_operator_t op, int64_t nsv, int64_t ncv, int64_t maxit, double tol,
              const double *init, svb_result_t *out);

/* ---- the step after the path: k-nearest neighbours in PCA space (neighbours.jl:19-86) ---------------------------- */
/* Replaces libcell `FindNeighbours{Euclidean,Cosine}{32,64}` (call sites neighbours.jl:39-41,50-52,61-63,72-74): X is the
 * n x d column-major coordinate matrix (dtype SVB_F32 | SVB_F64, column stride ldx = the reference's stride(X,2)), nn_index
 * n x k Int32 and distances n x k (same element type as X), both column-major and caller-owned, as in `ann!`. The search is
 * EXACT (brute force on the device) — what test/test_nn.jl compares the reference's approximate search against — so the
 * reference's `ntables` and `seed` arguments have no counterpart. Row i lists the k nearest cells of cell i by increasing
 * distance (ties: lower index first); include_self != 0 puts the cell itself first (distance 0), otherwise it is excluded.
 * Euclidean = sqrt(sum (x-y)^2), cosine = max(1 - x.y/(|x||y|), 0) (Distances.jl; a zero vector is at distance 1 from
 * everything). index_base 1 for Julia. Limits: k <= 64, d <= 128, n < 2^31. */
#define SVB_METRIC_EUCLIDEAN 0
#define SVB_METRIC_COSINE 1
int svb_knn(const void *X, int dtype, int64_t n, int64_t d, int64_t ldx, int64_t k, int metric,
            int include_self, int index_base, int32_t *nn_index, void *distances);
/* The same search on the coordinates Z = U*Diagonal(s) (embedding.jl:67) of a solve that are still in HBM, first `dims`
 * components (dims <= 0: all) — nearest_neighbours(em, k, dims=1:dims) without the round trip through the host. */
int svb_knn_result(svb_result_t r, int64_t dims, int64_t k, int metric, int include_self,
                   int index_base, int32_t *nn_index, double *distances);

/* ---- the consumer of the kNN graph: Jaccard index / shared nearest neighbours (neighbours.jl:88-131,263-270) -------- */
/* Replaces `_jaccard_index` (neighbours.jl:88-94 with a fixed k, :96-110 without): snn = nn' * nn — entry (i, j) = number
 * of neighbours cells i and j share — each stored x mapped to x / (k + (k - x)) in the output element type, then
 * `droptol!(snn, prune)` (entries with abs(x) <= prune removed). nn: the n x n neighbour graph as uploaded by
 * svb_csc_upload (column i = the neighbours of cell i, the layout `nearest_neighbours` returns, neighbours.jl:79; only the
 * PATTERN is read — upload the `true` entries; rows strictly ascending inside a column, as SparseMatrixCSC guarantees,
 * else SVB_EDIM). k > 0: the fixed neighbourhood size of `jaccard_index(nn, k)` / `shared_nearest_neighbours`; k <= 0: the
 * form without k, where column j uses diag(snn)[j] = its own number of neighbours (:99,103-106). dtype SVB_F32 | SVB_F64
 * = the reference's `dtype` (prune is rounded to it first, :128). out: n x n CSC handle with ascending rows and values of
 * that type (svb_matrix_download / svb_matrix_free). The sparse product is never formed: see csrc/snn.cu. */
int svb_jaccard_index(svb_matrix_t nn, int64_t k, double prune, int dtype, svb_matrix_t *out);

/* ---- synthetic count matrices (benchmark inputs; counter-based RNG, any shard reproducible) ---- */
/* Poisson counts x_ij ~ Poisson(L_i * p_j * f_{c(i),j}), L_i log-normal, p_j gamma-shaped,
 * K planted cell programs with fold-change `fold` on ~5% of genes each. Rows [row0,row1) of the
 * m_total x genes matrix are generated (a rank's shard). Values int32. */
int svb_synth_counts(int64_t m_total, int64_t genes, int64_t row0, int64_t row1,
                     double mean_nnz_per_cell, int64_t programs, double fold, uint64_t seed,
                     svb_matrix_t *out);
/* standard-normal vector from the same counter-based generator (IRLBA init). */
int svb_synth_normal(int64_t n, uint64_t seed, double *host_out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* SEVERO_B200_H */

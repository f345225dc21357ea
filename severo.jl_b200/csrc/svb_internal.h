// svb_internal.h — shared internals of libsevero_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/severo_b200.h"

namespace svb {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &msg);

// Every device allocation of the library goes through a block cache on top of cudaMalloc (core.cu): a plain
// cudaFree of the multi-GB Lanczos workspace costs ~100 ms per solve on a B200, a cached block costs nothing.
cudaError_t dev_malloc(void **p, size_t bytes);
cudaError_t dev_free(void *p);
void dev_release_cache();  // return every cached block to the driver

#ifndef SVB_NO_ALLOC_MACROS
#define cudaMalloc(pp, bytes) svb::dev_malloc((void **)(pp), (bytes))
#define cudaFree(p) svb::dev_free((void *)(p))
#endif

#define SVB_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            char _b[512];                                                                    \
            snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),     \
                     __FILE__, __LINE__, cudaGetErrorString(_e));                            \
            throw svb::Error(_e == cudaErrorMemoryAllocation ? SVB_ENOMEM : SVB_ECUDA, _b);  \
        }                                                                                    \
    } while (0)

#define SVB_CHECK(cond, code, msg)                                                           \
    do {                                                                                     \
        if (!(cond)) throw svb::Error((code), std::string(msg));                             \
    } while (0)


#define SVB_API_BEGIN try {
#define SVB_API_END                                   \
    }                                                 \
    catch (const svb::Error &e) {                     \
        svb::set_last_error(e.what());                \
        return e.code;                                \
    }                                                 \
    catch (const std::bad_alloc &) {                  \
        svb::set_last_error("host out of memory");    \
        return SVB_ENOMEM;                            \
    }                                                 \
    catch (const std::exception &e) {                 \
        svb::set_last_error(e.what());                \
        return SVB_EARG;                              \
    }                                                 \
    return SVB_OK;

#define SVB_LAUNCH_CHECK() SVB_CUDA(cudaGetLastError())

// ---- global context ---------------------------------------------------------------------------
struct Context {
    bool initialised = false;
    int device = -1;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool pool_ok = false;  // stream-ordered allocator available
    // profiling
    bool profile = false;
    double prof_ms[SVB_K_NCLASS] = {0};
    int64_t prof_launches[SVB_K_NCLASS] = {0};
    double prof_bytes[SVB_K_NCLASS] = {0};
    int64_t launches = 0;
    // comm
    int nranks = 1;
    int rank = 0;
    void *nccl_comm = nullptr;
    // per-context state owned by other translation units (opaque here): the block cache of the device allocator (core.cu),
    // the reduction scratch of the tall-skinny kernels (dense.cu), the peer-mailbox state (p2p.cu)
    void *alloc_cache = nullptr;
    void *dense_scratch = nullptr;
    void *p2p_state = nullptr;
};
// The context of the calling thread: the process-wide one (svb_init: one process drives one GPU), or — inside the worker
// threads of the in-process device group (multi.cu: svb_init_devices, one thread per GPU) — that worker's own.
Context &ctx();
void set_thread_context(Context *c);  // nullptr = back to the process-wide context
void require_init();

// RAII device buffer
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) SVB_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    T *take() { T *r = p; p = nullptr; n = 0; return r; }
};

// scoped timer for one kernel class: records events only when profiling is on.
struct KTimer {
    int cls;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double bytes;
    KTimer(int c, double algorithmic_bytes, int nlaunch = 1);
    ~KTimer();
};
void count_launch(int n = 1);

// ---- matrix / operator objects ------------------------------------------------------------------
}  // namespace svb

struct svb_matrix_s {
    int64_t nrow = 0, ncol = 0, nnz = 0;
    int vtype = SVB_F64;         // SVB_I32 | SVB_F32 | SVB_F64 on device
    int64_t *colptr = nullptr;   // [ncol+1], 0-based
    int32_t *rowidx = nullptr;   // [nnz], 0-based, ascending inside a column
    void *val = nullptr;         // [nnz]
    ~svb_matrix_s();
};

// count-level layouts of the scaled HVG operator (factored.cu): one 16-bit code per nonzero, the value is rebuilt from a
// per-cell level table and a per-gene scale
struct svb_operator_s;
struct svb_factored_s {
    int L = 0, log2L = 0;            // count levels 1..L are coded; everything else is an exception chunk (exact Float64 value)
    int log2R = 0;                   // cells per adjoint tile (R*L table entries in shared memory)
    int64_t Rc = 0;                  // cells a tile actually holds (<= R, a multiple of 16; whole rounds of the grid)
    int64_t R = 0, ntiles = 0;
    double *tlev = nullptr;          // [m*L] t_i[l] = log1p(sf*(l+1)/s_i), cell-major (forward)
    double *tlevA = nullptr;         // [ntiles*R*L] the same, level-major inside every adjoint tile
    double *inv = nullptr;           // [n] 1/sd
    int64_t *f_rowptr = nullptr;     // [m+1] forward: first chunk of a cell's row (chunk = 8 codes)
    void *f_code = nullptr;          // [f_chunks*8] u16 gene index (<< f_cshift: byte offsets when n <= 8190), pad = n
    int f_cshift = 0;
    uint8_t *f_meta = nullptr;       // [f_chunks] (level << 1) | last chunk of the row
    int64_t f_chunks = 0;
    int64_t *a_gptr = nullptr;       // [ntiles*n+1] adjoint: first chunk of a (tile, gene) segment
    void *a_code = nullptr;          // [a_chunks*8] u16 (level-1)*R + i_local, pad = R*L
    uint8_t *a_meta = nullptr;       // [a_chunks] last chunk of the segment
    int32_t *a_slices = nullptr;     // [ntiles*(K+1)] gene boundaries of the K per-warp slices of a tile
    unsigned int *counters = nullptr; // [2] tile counter / finished-CTA counter of the adjoint kernel (self-resetting)
    int64_t a_chunks = 0;
    double *partial = nullptr;       // [ntiles*(n+1)]
    int64_t nnz_main = 0, nnz_exc = 0;
    // exception entries (count > L, count < 1, clipped): NOT in the streams (round 2) — two small side matrices with the exact
    // Float64 value (times sd): the adjoint's reduce kernel and a per-cell kernel after the forward stream add them
    svb_matrix_s *exc = nullptr;     // gene-major (CSC): colptr[n+1], rowidx = cell, val f64 — the adjoint's view
    int64_t *e_segptr = nullptr;     // [n+1] the gene-major side matrix cut into segments of at most 4096 entries (balanced work: a few
                                     // dense genes hold most of the exceptions); e_segsum[e_nseg] = per-segment sums of value*w
    double *e_segsum = nullptr;
    int64_t e_nseg = 0;
    svb_matrix_s *excT = nullptr;    // cell-major (CSC of the transpose): colptr[m+1] over cells, rowidx = gene — the forward's view
    // bank-shifted replicas of the gathered tables (round 2): the builder picks, set by set, the replica of every entry so that
    // the 16 gathers of a half-warp fall in 16 different 8-byte banks (bipartite matching at build time, see factored.cu)
    int f_nrep = 1, f_stride = 0;    // forward: xs replica r starts at entry r*f_stride (f_stride = 5 mod 16); code = byte offset
    int a_nlr = 0, a_nrep = 1;       // adjoint: levels 1..a_nlr have a_nrep replicas, the others one copy
    int a_strideA = 0, a_levstride = 0, a_baseB = 0;  // entry of (l < a_nlr, r, i): l*a_levstride + r*a_strideA + i (a_levstride = 0 mod 16,
                                     // a_strideA = 5 mod 16: replica 0 keeps bank = cell mod 16); (l >= a_nlr, i): a_baseB + (l-a_nlr)*R + i
    int a_pad = 0, a_wbase = 0, a_tabsize = 0;  // pad entry (0.0), first of the R entries of w (exception chunks), table entries
    double f_passes = 0.0, a_passes = 0.0;      // average shared-memory passes per set of 16 gathers after the assignment (1 = no conflict)
    int fwd_grid = 0, adj_grid = 0;
    int64_t *fwd_ranges = nullptr, *fwd_rows = nullptr;  // [warps+1] first chunk / first row of every warp of the forward grid
    ~svb_factored_s();
};

struct svb_operator_s {
    int64_t m = 0, n = 0, nnz = 0;  // S is m x n (cells x genes)
    svb_factored_s *fact = nullptr; // count-level form (svb_operator_create_counts); the explicit layouts below are unused then
    bool dense = false;
    int vbytes = 8;                 // value storage width (8 = f64, 4 = f32)
    int ibytes = 2;                 // forward index width (2 = u16, 4 = i32)
    double *mu = nullptr;           // [n] or null
    // forward layout: CSR by cell
    int64_t *rowptr = nullptr;      // [m+1]
    void *fidx = nullptr;           // [nnz] u16 or i32 gene index
    void *fval = nullptr;           // [nnz]
    // adjoint layout: cells tiled by R, gene-major inside a tile
    int64_t R = 0, ntiles = 0;
    int log2R = 0;
    int fwd_lps = 0, adj_lps = 0, adj_gs = 1;
    int64_t *fwd_ranges = nullptr, *adj_ranges = nullptr;  // [grid+1] equal-nnz CTA boundaries
    int fwd_grid = 0, adj_grid = 0;          // persistent grid sizes (one resident wave), set at first launch  // launch shape knobs (0 = choose from the average segment length)
    int64_t *gptr = nullptr;        // [ntiles*n + 1] segment (tile, gene) -> offset
    uint16_t *rloc = nullptr;       // [nnz] cell index inside the tile
    void *aval = nullptr;           // [nnz]
    double *partial = nullptr;      // [ntiles * (n+1)] per-tile partial S'w (+ tile sum of w)
    // dense variant (column-major, ld = rows of the stored array)
    double *dA = nullptr;
    int64_t lda = 0;
    bool dense_transposed = false;
    // scratch
    double *xdev = nullptr, *ydev = nullptr;  // for svb_mul host staging
    double *tmp = nullptr;                    // [max(m,n)] product before the epilogue / allreduce
    double *scal = nullptr;                   // [8] device scalars
    ~svb_operator_s();
};

struct svb_result_s {
    int64_t m = 0, n = 0, nu = 0, iter = 0, mprod = 0;
    int info = 0;
    double *U = nullptr, *s = nullptr, *V = nullptr;  // device
    ~svb_result_s();
};

namespace svb {
// ---- scan.cu
void exclusive_scan_i64(int64_t *d_inout, int64_t n, cudaStream_t st);  // in place, n elements; returns nothing
// ---- matrix.cu
svb_matrix_s *matrix_alloc(int64_t nrow, int64_t ncol, int64_t nnz, int vtype);
size_t vtype_size(int vtype);
void csc_validate(const svb_matrix_s *a, const char *who);  // rows in range and strictly ascending inside every column (throws SVB_EDIM)
// ---- operator / spmv
void op_apply(svb_operator_s *op, bool trans, double alpha, const double *dx, double beta, double *dy,
              const double *axpy_coef_dev = nullptr, double axpy_sign = 0.0, const double *axpy_vec = nullptr);
int64_t *make_cta_ranges(const int64_t *ptr, int64_t nseg, int64_t nnz, int G);  // [G+1] equal-work CTA boundaries
// ---- factored.cu : products of the count-level form
void fact_fwd(svb_operator_s *op, double alpha, const double *dx, double beta, double *dy, const double *coef, double csign,
              const double *cvec);
void fact_adj_stage1(svb_operator_s *op, const double *dx);
double fact_fwd_bytes(const svb_operator_s *op);
double fact_adj_bytes(const svb_operator_s *op);
// ---- spmm.cu : k right-hand sides (scaling.jl:259-272), device pointers, column-major with leading dims
void op_apply_mm(svb_operator_s *op, bool trans, double alpha, const double *dX, int64_t ldx, double beta, double *dY, int64_t ldy,
                 int64_t k);
void op_apply_cols(svb_operator_s *op, bool trans, const double *dX, int64_t ldx, double *dY, int64_t ldy, int64_t k);  // any operator kind
void op_gram(svb_operator_s *op, double *G);  // G (n x n device, ld n) = S'S  (scaling.jl:274-296)
// ---- dense.cu (tall-skinny kernels; all pointers device)
// t[0..j) = X[:, 0..j)' * y      (X col-major L x j, leading dim ld)
struct P2PCtx;
void ts_gemv_t(const double *X, int64_t ld, int64_t L, int j, const double *y, double *t, int cls,
               const P2PCtx *produce_t = nullptr);
// y = beta*y + alpha * X[:, 0..j) * t ; nrm2_out (optional) receives sum(y.^2) of the result
void ts_gemv_n(const double *X, int64_t ld, int64_t L, int j, const double *t, double alpha, double beta,
               double *y, double *nrm2_out, int cls, const P2PCtx *consume_t = nullptr, const P2PCtx *produce_nrm = nullptr);
// out[:, 0..k) = X[:, 0..w) * P[0..w, 0..k)  (P device, col-major ldp), optional per-column scale
void ts_gemm(const double *X, int64_t ld, int64_t L, int w, const double *P, int ldp, int k, double *out,
             int64_t ldo, const double *colscale_dev);
void vec_sumsq(const double *x, int64_t L, double *out);            // out = sum x^2 (device scalar)
// y = x * (1/sqrt(*nrm2)) ; also writes sqrt(*nrm2) to *norm_out (device) and flags breakdown
bool vside_cgs_supported(int64_t n, int j);
void vside_cgs(const double *V, int64_t n, int j, double *f, double *nrm2, double *out, double *slot, int *flag, double eps);
void vec_normalize(const double *x, int64_t L, const double *nrm2_dev, double *y, double *norm_out,
                   int *flag_dev, double eps, const P2PCtx *consume_nrm = nullptr);
void vec_copy(const double *x, int64_t L, double *y);
void vec_fill_normal(double *x, int64_t L, uint64_t seed, uint64_t offset);
// ---- comm.cpp
void comm_allreduce_dev(double *dbuf, int64_t n);  // no-op when nranks == 1
// ---- p2p.cu : one-shot allreduce through IPC-mapped peer mailboxes (small messages)
void p2p_setup(int nranks, int rank);
void p2p_teardown();
bool p2p_ready();
bool p2p_allreduce(double *dbuf, int64_t n);  // false => not handled (use NCCL)
struct Mailbox;
Mailbox *p2p_local_alloc();                                              // in-process device group: phase 1 (per worker)
void p2p_local_connect(int nranks, int rank, Mailbox *const *all);       // phase 2 (per worker, after a host barrier)
int p2p_error();
// ---- bsvd.cu : device-side SVD of B + convergence test + restart bookkeeping (single CTA)
struct BsvdStatus {
    int converged;  // 1 converged, 0 not, -1 skipped because the breakdown flag was set
    int nconv;
    int k;
    int sweeps;
    double sigma0;
    double RF;
};
bool bsvd_supported(int w);
void bsvd_launch(int w, int nu, double *B, double *P, double *Q, double *sig, double *sig_prev, const double *nrm2F,
                 double *smax_io, double tol, double svtol, int k_in, const int *flag_dev, BsvdStatus *status_dev);
// ---- jacobi_svd.cpp : A (w x w, col-major) = P diag(s) Q', s descending
void small_svd(int w, const double *A, double *P, double *s, double *Q);
}  // namespace svb

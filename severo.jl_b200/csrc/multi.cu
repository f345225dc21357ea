// multi.cu — ONE process, ONE calling thread, N GPUs: the reference's boundary is a single synchronous call from one host
// thread (`ccall(("irlba", libcell), ...)`, src/irlba.jl:66-71), so a drop-in must reach all the GPUs of a node from that one
// call — no launcher, no rendezvous (SURVEY 8b "single process drives all GPUs", 8e ncclCommInitAll).
//
// svb_init_devices(ndev, devices) starts one WORKER THREAD per GPU. A worker owns a library context of its own (stream, block
// cache, reduction scratch, peer mailboxes: svb_internal.h Context) and runs the same single-device code as a rank of the
// one-process-per-GPU mode — the code is SPMD either way; only the plumbing differs:
//   * the communicator comes from ncclCommInitAll (no unique id to exchange);
//   * the peer mailboxes of the fused exchanges (p2p.cuh) are reached through plain peer access
//     (cudaDeviceEnablePeerAccess) instead of CUDA-IPC handles: all workers share the address space.
// The multi-device entry points take the caller's WHOLE SparseMatrixCSC (host arrays, Julia's 1-based Int64 indices accepted),
// shard it by cells inside — every worker cuts its own rows out of the host arrays in parallel and uploads them — run the solve
// on all devices, and write s, V (identical on all ranks) and the row blocks of U into the caller's buffers.
// Errors: the first failing worker's code and message are returned to the caller (svb_last_error); no exception crosses.
#include "svb_internal.h"

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

using namespace svb;

namespace svb {

void comm_init_all(int ndev, const int *devs, void **comms_out);  // comm.cpp
void comm_destroy_one(void *comm);

struct Worker {
    int rank = 0, device = 0;
    Context C;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, quit = false, done = false;
    int err_code = SVB_OK;
    std::string err_msg;
};

struct Group {
    std::vector<std::unique_ptr<Worker>> w;
    bool comm = false;
};
static Group *g_group = nullptr;

static void worker_main(Worker *W) {
    set_thread_context(&W->C);
    cudaSetDevice(W->device);
    for (;;) {
        std::function<void()> job;
        {
            std::unique_lock<std::mutex> lk(W->mu);
            W->cv.wait(lk, [&] { return W->has_job || W->quit; });
            if (W->quit) break;
            job = std::move(W->job);
            W->has_job = false;
        }
        int code = SVB_OK;
        std::string msg;
        try {
            job();
        } catch (const Error &e) {
            code = e.code;
            msg = e.what();
        } catch (const std::bad_alloc &) {
            code = SVB_ENOMEM;
            msg = "host out of memory";
        } catch (const std::exception &e) {
            code = SVB_EARG;
            msg = e.what();
        }
        {
            std::lock_guard<std::mutex> lk(W->mu);
            W->err_code = code;
            W->err_msg = msg;
            W->done = true;
        }
        W->cv.notify_all();
    }
    set_thread_context(nullptr);
}

// fn(rank) on every worker at once; returns when all have finished; throws the first error (lowest rank)
static void run_all(Group *G, const std::function<void(int)> &fn) {
    for (auto &W : G->w) {
        std::lock_guard<std::mutex> lk(W->mu);
        const int r = W->rank;
        W->job = [fn, r] { fn(r); };
        W->has_job = true;
        W->done = false;
        W->cv.notify_all();
    }
    for (auto &W : G->w) {
        std::unique_lock<std::mutex> lk(W->mu);
        W->cv.wait(lk, [&] { return W->done; });
    }
    for (auto &W : G->w)
        if (W->err_code != SVB_OK) throw Error(W->err_code, "device " + std::to_string(W->device) + ": " + W->err_msg);
}

// a public entry point called inside a worker: turn its return code into an exception carrying the worker thread's message
static void wcheck(int rc) {
    if (rc != SVB_OK) throw Error(rc, svb_last_error());
}

static Group *require_group() {
    if (!g_group) throw Error(SVB_ECUDA, "svb_init_devices has not been called (there is no CPU fallback)");
    return g_group;
}

// cells [lo, hi) of worker r: equal cell counts, boundaries on multiples of 4
static void shard_bounds(int64_t m, int n, std::vector<int64_t> &b) {
    b.assign((size_t)n + 1, 0);
    for (int r = 1; r < n; ++r) b[(size_t)r] = std::min<int64_t>(m, ((m * r / n) + 3) / 4 * 4);
    b[(size_t)n] = m;
    for (int r = 1; r <= n; ++r) b[(size_t)r] = std::max(b[(size_t)r], b[(size_t)r - 1]);
}

// the rows [lo, hi) of a host CSC (ascending rows inside a column) as a CSC of its own: 0-based int32 rows, values copied as they are
template <typename IdxT>
static void cut_rows(int64_t ncol, const int64_t *colptr, const IdxT *rowval, const char *nzval, size_t vsize, int base, int64_t lo,
                     int64_t hi, std::vector<int64_t> &cp, std::vector<int32_t> &rv, std::vector<char> &nz, std::vector<int64_t> &first) {
    cp.assign((size_t)ncol + 1, 0);
    first.assign((size_t)ncol, 0);
    for (int64_t j = 0; j < ncol; ++j) {
        const IdxT *b = rowval + (colptr[j] - base), *e = rowval + (colptr[j + 1] - base);
        const IdxT *p0 = std::lower_bound(b, e, (IdxT)(lo + base)), *p1 = std::lower_bound(p0, e, (IdxT)(hi + base));
        first[(size_t)j] = (colptr[j] - base) + (p0 - b);
        cp[(size_t)j + 1] = cp[(size_t)j] + (p1 - p0);
    }
    const int64_t nnz = cp[(size_t)ncol];
    rv.resize((size_t)std::max<int64_t>(nnz, 1));
    nz.resize((size_t)std::max<int64_t>(nnz, 1) * vsize);
    for (int64_t j = 0; j < ncol; ++j) {
        const int64_t cnt = cp[(size_t)j + 1] - cp[(size_t)j], src = first[(size_t)j], dst = cp[(size_t)j];
        for (int64_t k = 0; k < cnt; ++k) rv[(size_t)(dst + k)] = (int32_t)((int64_t)rowval[src + k] - base - lo);
        if (cnt) memcpy(nz.data() + (size_t)dst * vsize, nzval + (size_t)src * vsize, (size_t)cnt * vsize);
    }
}

static size_t host_vsize(int vtype) { return vtype == SVB_F64 || vtype == SVB_I64 ? 8 : 4; }

struct ShardedCSC {
    int64_t m, n;
    const int64_t *colptr;
    const void *rowval;
    int rowval_type;
    const void *nzval;
    int vtype, index_base;
};

// The piece of every column that worker r took: [first[j], first[j] + count[j]) of the caller's arrays
struct ShardPieces {
    std::vector<int64_t> first, count;
};

// upload worker r's cells; returns the device matrix (owned by the caller, to be freed on the same worker)
static svb_matrix_t upload_shard(const ShardedCSC &A, int64_t lo, int64_t hi, ShardPieces &pieces) {
    std::vector<int64_t> cp;
    std::vector<int32_t> rv;
    std::vector<char> nz;
    const size_t vs = host_vsize(A.vtype);
    if (A.rowval_type == SVB_I64)
        cut_rows<int64_t>(A.n, A.colptr, (const int64_t *)A.rowval, (const char *)A.nzval, vs, A.index_base, lo, hi, cp, rv, nz, pieces.first);
    else
        cut_rows<int32_t>(A.n, A.colptr, (const int32_t *)A.rowval, (const char *)A.nzval, vs, A.index_base, lo, hi, cp, rv, nz, pieces.first);
    pieces.count.resize((size_t)A.n);
    for (int64_t j = 0; j < A.n; ++j) pieces.count[(size_t)j] = cp[(size_t)j + 1] - cp[(size_t)j];
    svb_matrix_t h = nullptr;
    wcheck(svb_csc_upload(hi - lo, A.n, cp.data(), rv.data(), SVB_I32, nz.data(), A.vtype, 0, &h));  // validates the shard on the device
    return h;
}

// Phase 1 of the multi-device entry points: every worker cuts and uploads its cells — NO collective in here, so a malformed
// input fails on its worker and run_all reports it after all workers have returned (an error raised between two collectives
// would leave the peers waiting). The cut is a binary search per column and shard, which is only right when the rows ascend
// inside every column: each shard is validated on the device (ascending, in range), and here the pieces of a column must tile
// it exactly — together that IS "rows ascend strictly in the whole column". On failure the uploaded shards are released.
static void upload_all_shards(Group *G, const ShardedCSC &A, const std::vector<int64_t> &bounds, std::vector<svb_matrix_t> &mats,
                              const char *who) {
    const int N = (int)G->w.size();
    std::vector<ShardPieces> pieces((size_t)N);
    mats.assign((size_t)N, nullptr);
    auto release = [&] {
        try {
            run_all(G, [&](int r) {
                if (mats[(size_t)r]) svb_matrix_free(mats[(size_t)r]);
                mats[(size_t)r] = nullptr;
            });
        } catch (...) {
        }
    };
    try {
        run_all(G, [&](int r) { mats[(size_t)r] = upload_shard(A, bounds[(size_t)r], bounds[(size_t)r + 1], pieces[(size_t)r]); });
        for (int64_t j = 0; j < A.n; ++j) {
            int64_t at = A.colptr[j] - A.index_base;
            for (int r = 0; r < N; ++r) {
                SVB_CHECK(pieces[(size_t)r].first[(size_t)j] == at, SVB_EDIM, std::string(who) + ": row indices must ascend inside every column");
                at += pieces[(size_t)r].count[(size_t)j];
            }
            SVB_CHECK(at == A.colptr[j + 1] - A.index_base, SVB_EDIM, std::string(who) + ": row indices must ascend inside every column and lie in [0, m)");
        }
    } catch (...) {
        release();
        throw;
    }
}

static void check_csc_args(const ShardedCSC &A, const char *who) {
    SVB_CHECK(A.colptr && (A.rowval || A.colptr[A.n] == A.index_base), SVB_EARG, std::string(who) + ": null argument");
    SVB_CHECK(A.m >= 1 && A.n >= 1 && A.m < 2147483647LL, SVB_EDIM, std::string(who) + ": bad dimensions");
    SVB_CHECK(A.index_base == 0 || A.index_base == 1, SVB_EARG, std::string(who) + ": index_base must be 0 or 1");
    SVB_CHECK(A.rowval_type == SVB_I32 || A.rowval_type == SVB_I64, SVB_EARG, std::string(who) + ": rowval_type must be I32 or I64");
    SVB_CHECK(A.colptr[0] == A.index_base && A.colptr[A.n] >= A.index_base, SVB_EDIM, std::string(who) + ": malformed colptr");
}

// solve on every worker and scatter the result into the caller's buffers (U column-major m x nu, rows of worker r at lo_r)
static void solve_and_collect(Group *G, const std::vector<int64_t> &bounds, std::vector<svb_operator_t> &ops, int64_t m, int64_t n,
                              int64_t nu, int64_t m_b, int64_t maxit, double tol, double svtol, const double *init, double *s, double *U,
                              double *V, int64_t *iter, int64_t *mprod, int *info_out) {
    std::vector<int> infos(G->w.size(), 0);
    std::vector<int64_t> its(G->w.size(), 0), mps(G->w.size(), 0);
    run_all(G, [&](int r) {
        svb_result_t res = nullptr;
        wcheck(svb_irlba_solve(ops[(size_t)r], nu, m_b, maxit, 0, tol, svtol, init, nullptr, nullptr, nullptr, &res));
        struct Guard {
            svb_result_t p;
            ~Guard() { svb_result_free(p); }
        } guard{res};
        int64_t ml = 0;
        wcheck(svb_result_info(res, &ml, nullptr, nullptr, &its[(size_t)r], &mps[(size_t)r], &infos[(size_t)r]));
        cudaStream_t st = ctx().stream;
        const int64_t lo = bounds[(size_t)r];
        if (U && ml > 0)
            SVB_CUDA(cudaMemcpy2DAsync(U + lo, (size_t)m * 8, res->U, (size_t)ml * 8, (size_t)ml * 8, (size_t)nu, cudaMemcpyDeviceToHost, st));
        if (r == 0) {
            if (s) SVB_CUDA(cudaMemcpyAsync(s, res->s, (size_t)nu * 8, cudaMemcpyDeviceToHost, st));
            if (V) SVB_CUDA(cudaMemcpyAsync(V, res->V, (size_t)n * nu * 8, cudaMemcpyDeviceToHost, st));
        }
        SVB_CUDA(cudaStreamSynchronize(st));
    });
    if (iter) *iter = its[0];
    if (mprod) *mprod = mps[0];
    *info_out = infos[0];
}

static void free_ops(Group *G, std::vector<svb_operator_t> &ops) {
    try {
        run_all(G, [&](int r) {
            if (ops[(size_t)r]) svb_operator_free(ops[(size_t)r]);
            ops[(size_t)r] = nullptr;
        });
    } catch (...) {
    }
}

}  // namespace svb

extern "C" {

int svb_init_devices(int ndev, const int *devices) {
    SVB_API_BEGIN
    SVB_CHECK(g_group == nullptr, SVB_EARG, "svb_init_devices: the device group is already initialised");
    int have = 0;
    cudaError_t e = cudaGetDeviceCount(&have);
    if (e != cudaSuccess || have == 0)
        throw Error(SVB_ECUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libsevero_b200 has no CPU fallback");
    if (ndev <= 0) ndev = have;
    SVB_CHECK(ndev <= have && ndev <= 16, SVB_EARG, "svb_init_devices: more devices requested than present (or than 16)");
    std::vector<int> devs((size_t)ndev);
    for (int i = 0; i < ndev; ++i) {
        devs[(size_t)i] = devices ? devices[i] : i;
        SVB_CHECK(devs[(size_t)i] >= 0 && devs[(size_t)i] < have, SVB_EARG, "svb_init_devices: device index out of range");
        for (int k = 0; k < i; ++k) SVB_CHECK(devs[(size_t)k] != devs[(size_t)i], SVB_EARG, "svb_init_devices: a device is listed twice");
    }
    auto G = std::unique_ptr<Group>(new Group());
    for (int r = 0; r < ndev; ++r) {
        auto W = std::unique_ptr<Worker>(new Worker());
        W->rank = r;
        W->device = devs[(size_t)r];
        Worker *wp = W.get();
        W->th = std::thread(worker_main, wp);
        G->w.push_back(std::move(W));
    }
    try {
        // every worker: its own initialised context (the body of svb_init on that thread's context)
        run_all(G.get(), [&](int r) { wcheck(svb_init(devs[(size_t)r])); });
        if (ndev > 1) {
            std::vector<void *> comms((size_t)ndev, nullptr);
            comm_init_all(ndev, devs.data(), comms.data());
            G->comm = true;
            std::vector<Mailbox *> mail((size_t)ndev, nullptr);
            std::vector<int> peer_ok((size_t)ndev, 1);
            run_all(G.get(), [&](int r) {
                Context &C = ctx();
                C.nccl_comm = comms[(size_t)r];
                C.nranks = ndev;
                C.rank = r;
                for (int q = 0; q < ndev; ++q) {
                    if (q == r) continue;
                    int can = 0;
                    cudaDeviceCanAccessPeer(&can, devs[(size_t)r], devs[(size_t)q]);
                    if (!can) {
                        peer_ok[(size_t)r] = 0;
                        continue;
                    }
                    const cudaError_t pe = cudaDeviceEnablePeerAccess(devs[(size_t)q], 0);
                    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) peer_ok[(size_t)r] = 0;
                    cudaGetLastError();
                }
                mail[(size_t)r] = p2p_local_alloc();
            });
            bool all = true;
            for (int r = 0; r < ndev; ++r) all = all && peer_ok[(size_t)r] && mail[(size_t)r] != nullptr;
            if (all) run_all(G.get(), [&](int r) { p2p_local_connect(ndev, r, mail.data()); });
            else run_all(G.get(), [&](int) { p2p_teardown(); });  // no peer access somewhere: every exchange goes through NCCL
        }
    } catch (...) {
        for (auto &W : G->w) {
            {
                std::lock_guard<std::mutex> lk(W->mu);
                W->quit = true;
            }
            W->cv.notify_all();
            if (W->th.joinable()) W->th.join();
        }
        throw;
    }
    g_group = G.release();
    SVB_API_END
}

int svb_devices_info(int *ndev, int *devices, int *peer_mailboxes) {
    SVB_API_BEGIN
    Group *G = require_group();
    if (ndev) *ndev = (int)G->w.size();
    if (devices)
        for (size_t i = 0; i < G->w.size(); ++i) devices[i] = G->w[i]->device;
    if (peer_mailboxes) {
        std::vector<int> ready(G->w.size(), 0);
        run_all(G, [&](int r) { ready[(size_t)r] = p2p_ready() ? 1 : 0; });
        int all = G->w.size() > 1 ? 1 : 0;
        for (int v : ready) all = all && v;
        *peer_mailboxes = all;
    }
    SVB_API_END
}

int svb_shutdown_devices(void) {
    SVB_API_BEGIN
    if (!g_group) return SVB_OK;
    Group *G = g_group;
    try {
        run_all(G, [&](int) {
            Context &C = ctx();
            p2p_teardown();
            if (C.nccl_comm) {
                cudaStreamSynchronize(C.stream);
                comm_destroy_one(C.nccl_comm);
                C.nccl_comm = nullptr;
            }
            C.nranks = 1;
            C.rank = 0;
            svb_shutdown();
        });
    } catch (...) {
    }
    for (auto &W : G->w) {
        {
            std::lock_guard<std::mutex> lk(W->mu);
            W->quit = true;
        }
        W->cv.notify_all();
        if (W->th.joinable()) W->th.join();
    }
    delete G;
    g_group = nullptr;
    SVB_API_END
}

int svb_irlba_csc_devices(int64_t m, int64_t n, const int64_t *colptr, const void *rowval, int rowval_type, const void *nzval, int vtype,
                          int index_base, const double *mu, int64_t nu, int64_t m_b, int64_t maxit, double tol, double svtol,
                          const double *init, double *s, double *U, double *V, int64_t *iter, int64_t *mprod) {
    int info = SVB_OK;
    {
        SVB_API_BEGIN
        Group *G = require_group();
        const ShardedCSC A{m, n, colptr, rowval, rowval_type, nzval, vtype, index_base};
        check_csc_args(A, "svb_irlba_csc_devices");
        SVB_CHECK(init && s && U && V, SVB_EARG, "svb_irlba_csc_devices: null argument");
        SVB_CHECK(vtype == SVB_F64 || vtype == SVB_F32, SVB_EARG, "svb_irlba_csc_devices: Float64 / Float32 values (the scaled matrix)");
        if (!(svtol > 0.0)) svtol = tol;
        const int N = (int)G->w.size();
        std::vector<int64_t> bounds;
        shard_bounds(m, N, bounds);
        std::vector<svb_operator_t> ops((size_t)N, nullptr);
        std::vector<svb_matrix_t> mats;
        upload_all_shards(G, A, bounds, mats, "svb_irlba_csc_devices");
        try {
            run_all(G, [&](int r) {
                const int rc = svb_operator_create(mats[(size_t)r], mu, 0, &ops[(size_t)r]);
                svb_matrix_free(mats[(size_t)r]);
                mats[(size_t)r] = nullptr;
                wcheck(rc);
            });
            solve_and_collect(G, bounds, ops, m, n, nu, m_b, maxit, tol, svtol, init, s, U, V, iter, mprod, &info);
        } catch (...) {
            free_ops(G, ops);
            throw;
        }
        free_ops(G, ops);
        if (info == SVB_OK) return SVB_OK;
        SVB_API_END
    }
    svb::set_last_error(info == SVB_ENOCONV ? "irlba: not converged within maxit" : "irlba: starting vector in the null space");
    return info;
}

int svb_pca_counts_devices(int64_t m, int64_t n, const int64_t *colptr, const void *rowval, int rowval_type, const void *counts,
                           int vtype, int index_base, const int64_t *libsize, double scale_factor, double scale_max, int64_t nu,
                           int64_t m_b, int64_t maxit, double tol, double svtol, const double *init, double *mu_out, double *s,
                           double *U, double *V, int64_t *iter, int64_t *mprod) {
    int info = SVB_OK;
    {
        SVB_API_BEGIN
        Group *G = require_group();
        const ShardedCSC A{m, n, colptr, rowval, rowval_type, counts, vtype, index_base};
        check_csc_args(A, "svb_pca_counts_devices");
        SVB_CHECK(libsize && init && s && U && V, SVB_EARG, "svb_pca_counts_devices: null argument");
        SVB_CHECK(vtype == SVB_I32 || vtype == SVB_I64, SVB_EARG, "svb_pca_counts_devices: integer counts required");
        if (!(svtol > 0.0)) svtol = tol;
        const int N = (int)G->w.size();
        std::vector<int64_t> bounds;
        shard_bounds(m, N, bounds);
        std::vector<svb_operator_t> ops((size_t)N, nullptr);
        std::vector<double> mu0((size_t)n);
        std::vector<svb_matrix_t> mats;
        upload_all_shards(G, A, bounds, mats, "svb_pca_counts_devices");
        try {
            run_all(G, [&](int r) {
                const int64_t lo = bounds[(size_t)r];
                // moments = NULL: two parallel passes over the cells of ALL workers inside the build (allreduced)
                const int rc = svb_operator_create_counts(mats[(size_t)r], libsize + lo, scale_factor, nullptr, nullptr, scale_max, 0,
                                                          r == 0 ? mu0.data() : nullptr, &ops[(size_t)r]);
                svb_matrix_free(mats[(size_t)r]);
                mats[(size_t)r] = nullptr;
                wcheck(rc);
            });
            if (mu_out) memcpy(mu_out, mu0.data(), (size_t)n * 8);
            solve_and_collect(G, bounds, ops, m, n, nu, m_b, maxit, tol, svtol, init, s, U, V, iter, mprod, &info);
        } catch (...) {
            free_ops(G, ops);
            throw;
        }
        free_ops(G, ops);
        if (info == SVB_OK) return SVB_OK;
        SVB_API_END
    }
    svb::set_last_error(info == SVB_ENOCONV ? "irlba: not converged within maxit" : "irlba: starting vector in the null space");
    return info;
}

}  // extern "C"

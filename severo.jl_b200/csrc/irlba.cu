// irlba.cu — device-resident implicitly restarted Lanczos bidiagonalisation (IRLBA).
// Replaces the external `libcell.irlba` reached from src/irlba.jl:66-71 (and with it the Julia
// callbacks matmul/randv of src/irlba.jl:23-45: the operator products run on the GPU, nothing calls
// back into the host language). Algorithm: Baglama & Reichel's IRLBA as in B. W. Lewis' irlb.c, from
// which libcell derives (same signature / defaults: work = nu+7, tol = 1e-5, maxit = 1000); the loop
// is the one restated in SURVEY.md 8(c). libcell itself is un-vendored and un-pinned, so iterates are
// "parity unpinned"; converged results are pinned by test/test_irlba.jl's criteria.
//
// Execution model: V (n x w, replicated), W (m x w, cell-sharded), F, the bidiagonal entries and all
// norms stay in HBM. One Lanczos sweep is a host-sync-free sequence of kernels (norms are consumed
// from device scalars); the host reads the w diagonal / super-diagonal entries once per sweep, does
// the w x w SVD (small_svd.cpp) and the convergence test, and uploads the rotation for the restart
// product. A breakdown flag set by any normalisation makes the host redo that sweep in "careful"
// mode (norm checked on the host after every step, random restart vector from the device RNG).
#include "svb_internal.h"
#include "p2p.cuh"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstring>

using namespace svb;

svb_result_s::~svb_result_s() {
    if (U) cudaFree(U);
    if (s) cudaFree(s);
    if (V) cudaFree(V);
}

namespace svb {

static const double EPS23 = std::pow(2.220446049250313e-16, 2.0 / 3.0);

struct Solver {
    svb_operator_s *op;
    int64_t m, n;
    int w, nu;
    cudaStream_t st;
    DevBuf<double> V, V2, W, W2, F, T, Pd, Qd, sc;
    DevBuf<int> flag;
    bool careful = false;
    int64_t mprod = 0;
    uint64_t rng_calls = 0;

    double *Vc(int c) { return V.p + (int64_t)c * n; }
    double *Wc(int c) { return W.p + (int64_t)c * m; }
    double *nrm2W() { return sc.p + 0; }
    double *nrm2F() { return sc.p + 1; }

    double d2h(const double *p) {
        double v;
        SVB_CUDA(cudaMemcpyAsync(&v, p, sizeof(double), cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        return v;
    }

    // classical Gram-Schmidt of x (length L) against the first j columns of X; *nrm2 = |x|^2 after.
    // Cell-sharded vectors (W side): the coefficients and the norm are sums over the ranks. Fused path: the
    // producing kernels store into the peers' mailboxes and the consuming kernels wait + sum (p2p.cuh), so no
    // stand-alone collective is launched; `nrm_pending` tells finish() that the norm is still in the mailbox.
    bool nrm_pending = false;
    P2PCtx nrm_ctx{};
    void orthog(const double *X, int64_t L, int j, double *x, double *nrm2, bool sharded) {
        static const bool allow_fused = getenv("SVB_P2P_UNFUSED") == nullptr;
        const bool multi = sharded && ctx().nranks > 1;
        nrm_pending = false;
        if (j > 0) {
            P2PCtx pt{};
            const bool fuse = multi && allow_fused && !careful && j <= 256 && p2p_next_ctx(j, &pt);
            ts_gemv_t(X, L, L, j, x, T.p, SVB_K_REORTH, fuse ? &pt : nullptr);
            if (multi && !fuse) comm_allreduce_dev(T.p, j);
            const bool fuse_n = fuse && p2p_next_ctx(1, &nrm_ctx);
            ts_gemv_n(X, L, L, j, T.p, -1.0, 1.0, x, nrm2, SVB_K_REORTH, fuse ? &pt : nullptr, fuse_n ? &nrm_ctx : nullptr);
            if (fuse_n) {
                nrm_pending = true;
                return;
            }
        } else {
            vec_sumsq(x, L, nrm2);
        }
        if (multi) comm_allreduce_dev(nrm2, 1);
    }

    // x (|x|^2 in *nrm2) -> out = x/|x| ; |x| -> *slot. Returns false on an unrecoverable state.
    // careful mode: a tiny norm is replaced by a fresh random direction orthogonal to X[:, :j], slot = 0.
    void finish(const double *X, int64_t L, int j, double *x, double *nrm2, double *out, double *slot, bool sharded,
                bool *tiny) {
        if (tiny) *tiny = false;
        if (careful) {
            const double nrm = std::sqrt(d2h(nrm2));
            if (!(nrm >= EPS23)) {
                if (tiny) *tiny = true;
                const uint64_t off = ((uint64_t)(sharded ? ctx().rank : 0) << 40) + ((++rng_calls) << 48);
                vec_fill_normal(out, L, 0x5e7e70b200ull, off);
                double *tmpn = sc.p + 2;
                orthog(X, L, j, out, tmpn, sharded);
                vec_normalize(out, L, tmpn, out, nullptr, nullptr, 0.0);
                SVB_CUDA(cudaMemsetAsync(slot, 0, sizeof(double), st));
                return;
            }
        }
        vec_normalize(x, L, nrm2, out, slot, flag.p, EPS23, nrm_pending ? &nrm_ctx : nullptr);
        nrm_pending = false;
    }
};

// The one host read per sweep (32 bytes: converged, k, sweeps). The GPU is idle while the host reacts to it, so the read goes
// into a pinned buffer and the host SPINS on the stream: a copy into pageable memory waits inside the driver with a blocking wait
// whose wake-up showed up as 1 ms gaps per restart in some runs (0.129 instead of 0.118 s per C3 solve on identical kernels).
static void read_status(const BsvdStatus *dev, BsvdStatus *host, cudaStream_t st) {
    static thread_local BsvdStatus *pin = nullptr;  // one per host thread (the workers of the device group have their own)
    if (!pin) SVB_CUDA(cudaHostAlloc((void **)&pin, sizeof(BsvdStatus), cudaHostAllocDefault));
    SVB_CUDA(cudaMemcpyAsync(pin, dev, sizeof(BsvdStatus), cudaMemcpyDeviceToHost, st));
    cudaError_t e;
    while ((e = cudaStreamQuery(st)) == cudaErrorNotReady) {
    }
    SVB_CUDA(e);
    *host = *pin;
}

static void irlba_run(svb_operator_s *op, int64_t nu_, int64_t work_, int64_t maxit, int64_t restart, double tol, double svtol,
                      const double *init, const double *s0, const double *U0, const double *V0, svb_result_s *res) {
    Context &C = ctx();
    Solver S;
    S.op = op;
    S.m = op->m;
    S.n = op->n;
    S.st = C.stream;
    const int64_t m = S.m, n = S.n;
    const bool dbg = getenv("SVB_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    long long jacobi_sweeps = 0;
    double t_alloc = 0, t_svd = 0, t_wait = 0, t_issue = 0, t_start = now(), t_mark = 0;
    // global row count decides the work-size clamp (irlba.jl:56-58 uses min(m, n) of the whole matrix)
    double mglob = (double)m;
    if (C.nranks > 1) {
        DevBuf<double> d(1);
        SVB_CUDA(cudaMemcpyAsync(d.p, &mglob, 8, cudaMemcpyHostToDevice, S.st));
        comm_allreduce_dev(d.p, 1);
        SVB_CUDA(cudaMemcpyAsync(&mglob, d.p, 8, cudaMemcpyDeviceToHost, S.st));
        SVB_CUDA(cudaStreamSynchronize(S.st));
    }
    const int64_t minmn = std::min<int64_t>((int64_t)mglob, n);
    SVB_CHECK(nu_ >= 1 && nu_ <= minmn, SVB_EDIM, "irlba: nu must satisfy 1 <= nu <= min(m, n)");
    int64_t work = work_ > 0 ? work_ : nu_ + 7;
    if (work < nu_) work = nu_ + 1;
    if (work > minmn) work = minmn;
    SVB_CHECK(work >= nu_, SVB_EDIM, "irlba: work size smaller than nu");
    SVB_CHECK(restart >= 0 && restart < work, SVB_EDIM, "irlba: restart must be < work");
    SVB_CHECK(maxit >= 1, SVB_EDIM, "irlba: maxit must be >= 1");
    const int w = (int)work, nu = (int)nu_;
    S.w = w;
    S.nu = nu;
    t_mark = now();
    S.V.alloc((size_t)n * w);
    S.V2.alloc((size_t)n * w);
    S.W.alloc((size_t)m * w);
    S.W2.alloc((size_t)m * w);
    S.F.alloc((size_t)n);
    S.T.alloc((size_t)w + 8);
    S.Pd.alloc((size_t)w * w);
    S.Qd.alloc((size_t)w * w);
    S.sc.alloc(8);
    S.flag.alloc(1);
    t_alloc += now() - t_mark;
    SVB_CUDA(cudaMemsetAsync(S.flag.p, 0, sizeof(int), S.st));

    // B (w x w, column-major) lives on the device: the normalisation kernels write the diagonal / super-diagonal
    // entries straight into it, the fused SpMV epilogues read them back (s, r), bsvd_kernel decomposes it.
    DevBuf<double> Bm((size_t)w * w), sigd((size_t)w), sigprev((size_t)w), smaxd(1);
    DevBuf<BsvdStatus> statd(1);
    SVB_CUDA(cudaMemsetAsync(Bm.p, 0, (size_t)w * w * 8, S.st));
    SVB_CUDA(cudaMemsetAsync(sigprev.p, 0, (size_t)w * 8, S.st));
    SVB_CUDA(cudaMemsetAsync(smaxd.p, 0, 8, S.st));
    auto Bdiag = [&](int j) { return Bm.p + (size_t)j * w + j; };        // B[j, j]   = s_j
    auto Bsup = [&](int j) { return Bm.p + (size_t)(j + 1) * w + j; };   // B[j, j+1] = r_j
    const bool dev_svd = bsvd_supported(w) && getenv("SVB_HOST_SVD") == nullptr;
    std::vector<double> hB((size_t)w * w), P((size_t)w * w), Q((size_t)w * w), sig(w), sig_prev(w, 0.0), resid(w);
    int k = (int)restart;
    // start vector(s)
    SVB_CUDA(cudaMemcpyAsync(S.F.p, init, (size_t)n * 8, cudaMemcpyHostToDevice, S.st));
    vec_sumsq(S.F.p, n, S.nrm2F());
    if (k > 0) {
        SVB_CHECK(s0 && U0 && V0, SVB_EARG, "irlba: restart > 0 needs s, U, V inputs");
        SVB_CUDA(cudaMemcpyAsync(S.V.p, V0, (size_t)n * k * 8, cudaMemcpyHostToDevice, S.st));
        SVB_CUDA(cudaMemcpyAsync(S.W.p, U0, (size_t)m * k * 8, cudaMemcpyHostToDevice, S.st));
        std::fill(hB.begin(), hB.end(), 0.0);
        for (int i = 0; i < k; ++i) hB[(size_t)i * w + i] = s0[i];
        SVB_CUDA(cudaMemcpyAsync(Bm.p, hB.data(), (size_t)w * w * 8, cudaMemcpyHostToDevice, S.st));
        SVB_CUDA(cudaStreamSynchronize(S.st));
    }
    // warm restart (irlba.jl:87-99): the new start vector must be orthogonal to the supplied right vectors; the
    // reference passes the raw random `init` (its own test of this path is @test_broken, test_irlba.jl:60-62)
    if (k > 0) S.orthog(S.V.p, n, k, S.F.p, S.nrm2F(), false);
    vec_normalize(S.F.p, n, S.nrm2F(), S.Vc(k), nullptr, nullptr, 0.0);

    double smax = 0.0;
    int64_t iter = 0;
    int info = SVB_ENOCONV;
    bool have_svd = false;
    while (iter < maxit) {
        int j = (iter > 0 || restart > 0) ? k : 0;
        t_mark = now();
        bool tiny = false;
        // W_j = S*V_j ; orthogonalise against the kept W ; normalise
        op_apply(op, false, 1.0, S.Vc(j), 0.0, S.Wc(j));
        S.mprod++;
        S.orthog(S.W.p, m, j, S.Wc(j), S.nrm2W(), true);
        S.finish(S.W.p, m, j, S.Wc(j), S.nrm2W(), S.Wc(j), Bdiag(j), true, &tiny);
        if (tiny && iter == 0 && j == 0) {
            info = SVB_ENULLSPACE;
            break;
        }
        while (j < w) {
            // F = S'*W_j - s*V_j ; orthogonalise against V[:, :j+1]
            op_apply(op, true, 1.0, S.Wc(j), 0.0, S.F.p, Bdiag(j), -1.0, S.Vc(j));
            S.mprod++;
            // (gene side: whole on every rank — one cluster launch does coefficients, update, norm and the next basis vector)
            const bool vfused = !S.careful && vside_cgs_supported(n, j + 1);
            if (vfused) vside_cgs(S.V.p, n, j + 1, S.F.p, S.nrm2F(), j + 1 < w ? S.Vc(j + 1) : nullptr, j + 1 < w ? Bsup(j) : nullptr, S.flag.p, EPS23);
            else S.orthog(S.V.p, n, j + 1, S.F.p, S.nrm2F(), false);
            if (j + 1 < w) {
                if (!vfused) S.finish(S.V.p, n, j + 1, S.F.p, S.nrm2F(), S.Vc(j + 1), Bsup(j), false, nullptr);
                // W_{j+1} = S*V_{j+1} - r*W_j ; orthogonalise against W[:, :j+1]
                op_apply(op, false, 1.0, S.Vc(j + 1), 0.0, S.Wc(j + 1), Bsup(j), -1.0, S.Wc(j));
                S.mprod++;
                S.orthog(S.W.p, m, j + 1, S.Wc(j + 1), S.nrm2W(), true);
                S.finish(S.W.p, m, j + 1, S.Wc(j + 1), S.nrm2W(), S.Wc(j + 1), Bdiag(j + 1), true, nullptr);
            }
            ++j;
        }
        // ---- end of sweep: SVD of B, convergence test, restart size (device), one small host read ----
        int hflag = 0;
        BsvdStatus hs{};
        bool converged = false;
        if (dev_svd) {
            bsvd_launch(w, nu, Bm.p, S.Pd.p, S.Qd.p, sigd.p, sigprev.p, S.nrm2F(), smaxd.p, tol, svtol, k, S.flag.p, statd.p);
            t_issue += now() - t_mark;
            t_mark = now();
            read_status(statd.p, &hs, S.st);
            t_wait += now() - t_mark;
            hflag = hs.converged < 0;
        } else {
            double nF2 = 0.0;
            SVB_CUDA(cudaMemcpyAsync(hB.data(), Bm.p, (size_t)w * w * 8, cudaMemcpyDeviceToHost, S.st));
            SVB_CUDA(cudaMemcpyAsync(&nF2, S.nrm2F(), 8, cudaMemcpyDeviceToHost, S.st));
            SVB_CUDA(cudaMemcpyAsync(&hflag, S.flag.p, sizeof(int), cudaMemcpyDeviceToHost, S.st));
            t_issue += now() - t_mark;
            t_mark = now();
            SVB_CUDA(cudaStreamSynchronize(S.st));
            t_wait += now() - t_mark;
            if (!hflag || S.careful) {
                t_mark = now();
                small_svd(w, hB.data(), P.data(), sig.data(), Q.data());
                t_svd += now() - t_mark;
                const double RF = std::sqrt(nF2);
                for (int i = 0; i < w; ++i) resid[i] = RF * P[(size_t)i * w + (w - 1)];
                smax = std::max(smax, sig[0]);
                int nconv = 0;
                for (int i = 0; i < nu; ++i) {  // the nu wanted Ritz values only (convergence test)
                    const double ratio = std::fabs(sig_prev[i] - sig[i]) / sig[i];
                    if (std::fabs(resid[i]) < tol * smax && ratio < svtol) ++nconv;
                }
                hs.converged = (nconv >= nu || hB[(size_t)(w - 1) * w + (w - 1)] == 0.0 || RF <= 1000.0 * 2.220446049250313e-16 * smax) ? 1 : 0;
                hs.nconv = nconv;
                int kk = std::max(k, nu + nconv);
                kk = std::min(kk, w - 3);
                hs.k = std::max(kk, 1);
                SVB_CUDA(cudaMemcpyAsync(S.Pd.p, P.data(), (size_t)w * w * 8, cudaMemcpyHostToDevice, S.st));
                SVB_CUDA(cudaMemcpyAsync(S.Qd.p, Q.data(), (size_t)w * w * 8, cudaMemcpyHostToDevice, S.st));
                SVB_CUDA(cudaMemcpyAsync(sigd.p, sig.data(), (size_t)w * 8, cudaMemcpyHostToDevice, S.st));
                if (!hs.converged) {
                    sig_prev = sig;
                    std::fill(hB.begin(), hB.end(), 0.0);
                    for (int i = 0; i < hs.k; ++i) {
                        hB[(size_t)i * w + i] = sig[i];
                        hB[(size_t)hs.k * w + i] = resid[i];
                    }
                    SVB_CUDA(cudaMemcpyAsync(Bm.p, hB.data(), (size_t)w * w * 8, cudaMemcpyHostToDevice, S.st));
                }
                SVB_CUDA(cudaStreamSynchronize(S.st));
            }
        }
        if (C.nranks > 1 && p2p_error()) throw Error(SVB_ENCCL, "peer-memory allreduce timed out (a rank is missing)");
        if (hflag && !S.careful) {
            // a (near) breakdown happened somewhere in this sweep: redo it with host-checked norms
            S.careful = true;
            SVB_CUDA(cudaMemsetAsync(S.flag.p, 0, sizeof(int), S.st));
            continue;
        }
        if (hflag) SVB_CUDA(cudaMemsetAsync(S.flag.p, 0, sizeof(int), S.st));
        if (hflag && dev_svd) {
            // careful mode already replaced every tiny vector; the flag is stale (set by the fast normalise of a
            // legitimately tiny-but-accepted norm): run the decomposition with the flag cleared
            bsvd_launch(w, nu, Bm.p, S.Pd.p, S.Qd.p, sigd.p, sigprev.p, S.nrm2F(), smaxd.p, tol, svtol, k, S.flag.p, statd.p);
            read_status(statd.p, &hs, S.st);
        }
        have_svd = true;
        jacobi_sweeps += hs.sweeps;
        converged = hs.converged == 1;
        ++iter;
        if (converged) {
            info = SVB_OK;
            break;
        }
        if (iter >= maxit) break;
        k = hs.k;
        // restart: V[:, :k] = V*Q[:, :k] ; V[:, k] = F/|F| ; W[:, :k] = W*P[:, :k] ; B = [diag(sig) | resid] (already written)
        ts_gemm(S.V.p, n, n, w, S.Qd.p, w, k, S.V2.p, n, nullptr);
        vec_normalize(S.F.p, n, S.nrm2F(), S.V2.p + (int64_t)k * n, nullptr, nullptr, 0.0);
        ts_gemm(S.W.p, m, m, w, S.Pd.p, w, k, S.W2.p, m, nullptr);
        std::swap(S.V, S.V2);
        std::swap(S.W, S.W2);
    }
    res->m = m;
    res->n = n;
    res->nu = nu;
    res->iter = iter;
    res->mprod = S.mprod;
    res->info = info;
    t_mark = now();
    SVB_CUDA(cudaMalloc((void **)&res->U, (size_t)m * nu * 8));
    SVB_CUDA(cudaMalloc((void **)&res->V, (size_t)n * nu * 8));
    SVB_CUDA(cudaMalloc((void **)&res->s, (size_t)nu * 8));
    t_alloc += now() - t_mark;
    if (have_svd) {
        SVB_CUDA(cudaMemcpyAsync(res->s, sigd.p, (size_t)nu * 8, cudaMemcpyDeviceToDevice, S.st));
        ts_gemm(S.W.p, m, m, w, S.Pd.p, w, nu, res->U, m, nullptr);
        ts_gemm(S.V.p, n, n, w, S.Qd.p, w, nu, res->V, n, nullptr);
    } else {
        SVB_CUDA(cudaMemsetAsync(res->U, 0, (size_t)m * nu * 8, S.st));
        SVB_CUDA(cudaMemsetAsync(res->V, 0, (size_t)n * nu * 8, S.st));
        SVB_CUDA(cudaMemsetAsync(res->s, 0, (size_t)nu * 8, S.st));
    }
    SVB_CUDA(cudaStreamSynchronize(S.st));
    if (dbg)
        fprintf(stderr, "[svb irlba] total %.2f ms: alloc %.2f, issue %.2f, wait %.2f, svd %.2f (iters %lld, mprod %lld, jacobi sweeps %lld)\n",
                (now() - t_start) * 1e3, t_alloc * 1e3, t_issue * 1e3, t_wait * 1e3, t_svd * 1e3, (long long)iter, (long long)S.mprod, jacobi_sweeps);
}

__global__ void colscale_kernel(const double *__restrict__ U, const double *__restrict__ s, int64_t m, int64_t nu,
                                double *__restrict__ out) {
    const int64_t total = m * nu;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = U[i] * s[i / m];
}

__global__ void tssvd_sigma_kernel(const double *__restrict__ lambda, int64_t nu, double *__restrict__ sigma, double *__restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nu) {
        const double s = sqrt(fmax(lambda[i], 0.0));
        sigma[i] = s;
        inv[i] = s > 0.0 ? 1.0 / s : 0.0;
    }
}

// tssvd (embedding.jl:30-44): C = Hermitian(A'A); (lambda, phi) = the nsv largest eigenpairs of C; Sigma = sqrt(lambda);
// U = A*phi*inv(Diagonal(Sigma)); SVD(U, Sigma, phi'). The reference hands C to Arpack's `eigs` on the host; here C is
// built on the device (op_gram), stays there, and its eigenpairs come from the same device IRLBA applied to the dense
// symmetric positive semi-definite C (singular triplets of C = its eigenpairs). With cells sharded every rank holds the
// same allreduced C and solves the n x n problem redundantly (no exchange: nranks is masked for that solve).
static void tssvd_run(svb_operator_s *op, int64_t nu, int64_t work, int64_t maxit, double tol, const double *init, svb_result_s *res) {
    Context &C = ctx();
    cudaStream_t st = C.stream;
    const int64_t n = op->n, m = op->m;
    SVB_CHECK(nu >= 1 && nu <= n, SVB_EDIM, "tssvd: nsv must satisfy 1 <= nsv <= n");
    svb_operator_s gop;  // dense n x n operator over C; its destructor releases the buffers
    gop.dense = true;
    gop.m = n;
    gop.n = n;
    gop.nnz = n * n;
    gop.lda = n;
    gop.ibytes = 0;
    SVB_CUDA(cudaMalloc((void **)&gop.dA, (size_t)n * n * sizeof(double)));
    SVB_CUDA(cudaMalloc((void **)&gop.tmp, (size_t)n * sizeof(double)));
    SVB_CUDA(cudaMalloc((void **)&gop.scal, 8 * sizeof(double)));
    op_gram(op, gop.dA);
    svb_result_s eig;
    {
        struct Replicated {
            int saved;
            Replicated() : saved(ctx().nranks) { ctx().nranks = 1; }
            ~Replicated() { ctx().nranks = saved; }
        } guard;
        irlba_run(&gop, nu, work, maxit, 0, tol, tol, init, nullptr, nullptr, nullptr, &eig);
    }
    res->m = m;
    res->n = n;
    res->nu = nu;
    res->iter = eig.iter;
    res->mprod = eig.mprod;
    res->info = eig.info;
    SVB_CUDA(cudaMalloc((void **)&res->U, (size_t)m * nu * 8));
    SVB_CUDA(cudaMalloc((void **)&res->s, (size_t)nu * 8));
    res->V = eig.V;  // phi
    eig.V = nullptr;
    DevBuf<double> inv((size_t)nu), AV((size_t)m * nu);
    tssvd_sigma_kernel<<<(unsigned)((nu + 127) / 128), 128, 0, st>>>(eig.s, nu, res->s, inv.p);
    count_launch();
    SVB_LAUNCH_CHECK();
    op_apply_cols(op, false, res->V, n, AV.p, m, nu);
    colscale_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((m * nu + 255) / 256, 148 * 8)), 256, 0, st>>>(AV.p, inv.p, m, nu, res->U);
    count_launch();
    SVB_LAUNCH_CHECK();
    SVB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace svb

extern "C" {

int svb_irlba_solve(svb_operator_t op, int64_t nu, int64_t m_b, int64_t maxit, int64_t restart, double tol, double svtol,
                    const double *init, const double *s0, const double *U0, const double *V0, svb_result_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(op && init && out, SVB_EARG, "svb_irlba_solve: null argument");
    if (!(svtol > 0.0)) svtol = tol;
    auto *r = new svb_result_s();
    try {
        irlba_run(op, nu, m_b, maxit, restart, tol, svtol, init, s0, U0, V0, r);
    } catch (...) {
        delete r;
        throw;
    }
    *out = r;
    SVB_API_END
}

int svb_result_info(svb_result_t r, int64_t *m, int64_t *n, int64_t *nu, int64_t *iter, int64_t *mprod, int *info) {
    SVB_API_BEGIN
    SVB_CHECK(r, SVB_EARG, "null result handle");
    if (m) *m = r->m;
    if (n) *n = r->n;
    if (nu) *nu = r->nu;
    if (iter) *iter = r->iter;
    if (mprod) *mprod = r->mprod;
    if (info) *info = r->info;
    SVB_API_END
}

int svb_result_download(svb_result_t r, double *s, double *U, double *V, int scale_u) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(r, SVB_EARG, "null result handle");
    cudaStream_t st = ctx().stream;
    if (s) SVB_CUDA(cudaMemcpyAsync(s, r->s, (size_t)r->nu * 8, cudaMemcpyDeviceToHost, st));
    if (V) SVB_CUDA(cudaMemcpyAsync(V, r->V, (size_t)r->n * r->nu * 8, cudaMemcpyDeviceToHost, st));
    if (U) {
        if (scale_u) {
            // coordinates Z = U * Diagonal(s)  (embedding.jl:67)
            DevBuf<double> z((size_t)r->m * r->nu);
            colscale_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((r->m * r->nu + 255) / 256, 148 * 8)), 256, 0, st>>>(
                r->U, r->s, r->m, r->nu, z.p);
            count_launch();
            SVB_LAUNCH_CHECK();
            SVB_CUDA(cudaMemcpyAsync(U, z.p, (size_t)r->m * r->nu * 8, cudaMemcpyDeviceToHost, st));
            SVB_CUDA(cudaStreamSynchronize(st));
        } else {
            SVB_CUDA(cudaMemcpyAsync(U, r->U, (size_t)r->m * r->nu * 8, cudaMemcpyDeviceToHost, st));
        }
    }
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_API_END
}

int svb_result_free(svb_result_t r) {
    SVB_API_BEGIN
    if (r) {
        if (ctx().initialised) cudaStreamSynchronize(ctx().stream);
        delete r;
    }
    SVB_API_END
}

int svb_irlba(svb_operator_t op, int64_t nu, int64_t m_b, int64_t maxit, int64_t restart, double tol, double svtol,
              const double *init, double *s, double *U, double *V, int64_t *iter, int64_t *mprod) {
    svb_result_t r = nullptr;
    int rc = svb_irlba_solve(op, nu, m_b, maxit, restart, tol, svtol, init, s, U, V, &r);
    if (rc != SVB_OK) return rc;
    if (iter) *iter = r->iter;
    if (mprod) *mprod = r->mprod;
    const int info = r->info;
    rc = svb_result_download(r, s, U, V, 0);
    svb_result_free(r);
    if (rc != SVB_OK) return rc;
    if (info != SVB_OK) svb::set_last_error(info == SVB_ENOCONV ? "irlba: not converged within maxit" : "irlba: starting vector in the null space");
    return info;
}

int svb_tssvd(svb_operator_t op, int64_t nsv, int64_t ncv, int64_t maxit, double tol, const double *init, svb_result_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(op && init && out, SVB_EARG, "svb_tssvd: null argument");
    // eigs' tol = 0.0 (embedding.jl:30) asks for machine precision; the restart test here also compares successive Ritz
    // values relatively (svtol = tol), and those carry eps*lambda_max of rounding noise: 1e-12 leaves lambda_1/lambda_nsv up
    // to ~5e3 of room and gives sqrt(lambda) to ~1e-15 (residual^2/gap).
    if (!(tol > 0.0)) tol = 1e-12;
    auto *r = new svb_result_s();
    try {
        tssvd_run(op, nsv, ncv, maxit, tol, init, r);
    } catch (...) {
        delete r;
        throw;
    }
    *out = r;
    SVB_API_END
}

}  // extern "C"

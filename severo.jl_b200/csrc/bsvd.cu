// bsvd.cu — the per-restart bookkeeping of IRLBA on the device, in ONE single-CTA kernel:
//   1. SVD of the w x w projected matrix B (one-sided Jacobi, parallel round-robin ordering, G and the
//      accumulated right rotations in shared memory, one warp per column pair);
//   2. Ritz residuals  res_i = |F| * P[w-1, i], S_max, convergence count over the nu wanted values;
//   3. the restart size k and the next B = [diag(sigma_1..k) | res column] written back in place.
// The host reads one small status struct per sweep (it needs k and `converged` to launch the restart
// products); P and Q never leave the device. libcell does this on the host with LAPACK dgesdd (call site
// src/irlba.jl:66-71); with the cells sharded over 8 GPUs a 2 ms host SVD + uploads per sweep is >15 % of
// the solve, the kernel takes well under 0.1 ms. All ranks run it redundantly on identical inputs.
#include "svb_internal.h"

#include <algorithm>

namespace svb {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// LP = lanes per column pair: 32 (a whole warp per pair, one pair per warp) when a round has at most 32 pairs (w <= 64), else 16
// (two pairs per warp). ncu on the round-1 version (always 16, 15 warps at w = 57: profiles/r04_bsvd.md) shows a pure latency
// chain — 380 dependent-ish instructions per warp and round at ~9.5 cycles each, issue slots 40 % busy with 4 warps per scheduler:
// a whole warp per pair halves the trip counts of the dot / rotation loops and doubles the warps that hide each other's latency.
template <int LP>
__global__ void __launch_bounds__(1024) bsvd_kernel(int w, int nu, double *__restrict__ B, double *__restrict__ P,
                                                    double *__restrict__ Q, double *__restrict__ sig,
                                                    double *__restrict__ sig_prev, const double *__restrict__ nrm2F,
                                                    double *__restrict__ smax_io, double tol, double svtol, int k_in,
                                                    const int *__restrict__ flag, BsvdStatus *__restrict__ status) {
    extern __shared__ double sm[];
    if (*flag) {  // a normalisation in this sweep hit a (near) breakdown: the host redoes the sweep, B untouched
        if (threadIdx.x == 0) status->converged = -1;
        return;
    }
    const int ld = w | 1;
    double *G = sm;                      // [w][ld] column-major: column c at G + c*ld
    double *V = G + (size_t)w * ld;      // accumulated right rotations
    double *nrm = V + (size_t)w * ld;    // [w]
    double *res = nrm + w;               // [w]
    __shared__ int rotated, coarse;
    __shared__ int rank_of[256];
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    for (int idx = tid; idx < w * w; idx += nthreads) {
        const int c = idx / w, i = idx - c * w;
        G[c * ld + i] = B[idx];
        V[c * ld + i] = (i == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    const double eps = 2.220446049250313e-16;
    const int n = (w + 1) & ~1;  // even number of players (index w is a bye when w is odd)
    int sweeps = 0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        if (tid == 0) {
            rotated = 0;
            coarse = 0;
        }
        for (int c = warp; c < w; c += nwarps) {
            double a = 0.0;
            for (int i = lane; i < w; i += 32) a = fma(G[c * ld + i], G[c * ld + i], a);
            a = warp_sum_d(a);
            if (lane == 0) nrm[c] = a;
        }
        __syncthreads();
        for (int r = 0; r < n - 1; ++r) {
            // one column pair per group of LP lanes: all n/2 pairs of a round run concurrently
            constexpr int PPW = 32 / LP;  // pairs per warp
            const int sub = lane / LP;
            for (int pi = PPW * warp + sub; pi - sub < n / 2; pi += PPW * nwarps) {
                const int hl = lane % LP;
                int p, q;
                if (pi == 0) {
                    p = n - 1;
                    q = r;
                } else {
                    p = (r + pi) % (n - 1);
                    q = (r - pi + (n - 1)) % (n - 1);
                }
                if (p > q) { const int t = p; p = q; q = t; }
                const bool live = (pi < n / 2) && (q < w);
                double *gp = G + (live ? p : 0) * ld, *gq = G + (live ? q : 0) * ld;
                double g = 0.0;
                if (live)
                    for (int i = hl; i < w; i += LP) g = fma(gp[i], gq[i], g);
#pragma unroll
                for (int o = LP / 2; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);  // stays inside the lane group
                // squared column norms are cached (exact at the start of every sweep, updated by the rotation
                // formulas in between): one reduction per pair instead of three
                const double a = live ? nrm[p] : 0.0, b = live ? nrm[q] : 0.0;
                if (live && g * g > (eps * eps) * (a * b)) {  // |g| > eps*sqrt(a*b) without the sqrt
                    // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (b-a)/(2g) == 2g*sign(d) / (|d| + hypot(d, 2g)).
                    // fp64 sqrt / divide are ~500-cycle software sequences on the critical path of a round: seed with
                    // the fp32 SFU and polish with Newton steps in fp64 (c, s come from the same t, so c^2 + s^2 = 1 to
                    // rounding whatever the last bits of t are).
                    const double d = b - a, g2 = 2.0 * g;
                    const double q2 = fma(d, d, g2 * g2);
                    double t, c;
                    if (q2 > 1e-30 && q2 < 1e30) {
                        double y = (double)rsqrtf((float)q2);
                        y = y * fma(-0.5 * q2, y * y, 1.5);
                        y = y * fma(-0.5 * q2, y * y, 1.5);
                        const double x = fabs(d) + q2 * y;         // |d| + sqrt(q2)
                        double rr = (double)__frcp_rn((float)x);
                        rr = rr * fma(-x, rr, 2.0);
                        rr = rr * fma(-x, rr, 2.0);                // 1/x
                        t = (d >= 0.0 ? g2 : -g2) * rr;
                        const double u = fma(t, t, 1.0);
                        c = (double)rsqrtf((float)u);
                        c = c * fma(-0.5 * u, c * c, 1.5);
                        c = c * fma(-0.5 * u, c * c, 1.5);
                    } else {  // outside the fp32 seed range: plain fp64 sqrt / divide
                        t = (d >= 0.0 ? g2 : -g2) / (fabs(d) + sqrt(q2));
                        c = rsqrt(fma(t, t, 1.0));
                    }
                    const double sn = c * t;
                    __syncwarp(LP == 32 ? 0xffffffffu : (0xffffu << (lane & 16)));  // every lane of the group has read nrm[p], nrm[q]
                    if (hl == 0) {
                        nrm[p] = fmax(0.0, a - t * g);
                        nrm[q] = fmax(0.0, b + t * g);
                        rotated = 1;
                        if (g * g > 1e-16 * (a * b)) coarse = 1;  // a pair still above sqrt(eps)-level coupling
                    }
                    double *vp = V + p * ld, *vq = V + q * ld;
                    for (int i = hl; i < w; i += LP) {
                        const double x = gp[i], y = gq[i];
                        gp[i] = c * x - sn * y;
                        gq[i] = sn * x + c * y;
                        const double vx = vp[i], vy = vq[i];
                        vp[i] = c * vx - sn * vy;
                        vq[i] = sn * vx + c * vy;
                    }
                }
            }
            __syncthreads();
        }
        ++sweeps;
        // Jacobi converges quadratically: a sweep in which every coupling met was below 1e-8 leaves them at ~1e-16, i.e. at
        // rounding level — no further (verification) sweep is needed; `rotated` alone would always cost one sweep of no-ops
        const int any = rotated && coarse;
        __syncthreads();
        if (!any) break;
    }
    // singular values = column norms; rank them (descending, ties by column index)
    for (int c = warp; c < w; c += nwarps) {
        double a = 0.0;
        for (int i = lane; i < w; i += 32) a = fma(G[c * ld + i], G[c * ld + i], a);
        a = warp_sum_d(a);
        if (lane == 0) nrm[c] = sqrt(a);
    }
    __syncthreads();
    for (int c = tid; c < w; c += nthreads) {
        int rk = 0;
        const double mine = nrm[c];
        for (int o = 0; o < w; ++o) rk += (nrm[o] > mine) || (nrm[o] == mine && o < c);
        rank_of[c] = rk;
    }
    __syncthreads();
    for (int idx = tid; idx < w * w; idx += nthreads) {
        const int c = idx / w, i = idx - c * w;
        const int k = rank_of[c];
        const double inv = nrm[c] > 0.0 ? 1.0 / nrm[c] : 0.0;
        P[(size_t)k * w + i] = G[c * ld + i] * inv;
        Q[(size_t)k * w + i] = V[c * ld + i];
    }
    const double last_diag = B[(size_t)(w - 1) * w + (w - 1)];  // S of the last Lanczos step (0 => invariant subspace)
    const double RF = sqrt(*nrm2F);
    for (int c = tid; c < w; c += nthreads) {
        const int k = rank_of[c];
        const double inv = nrm[c] > 0.0 ? 1.0 / nrm[c] : 0.0;
        sig[k] = nrm[c];
        res[k] = RF * (G[c * ld + (w - 1)] * inv);  // RF * P[w-1, k]
    }
    __syncthreads();
    __shared__ int sh_k, sh_conv;
    if (tid == 0) {
        const double smax = fmax(*smax_io, sig[0]);
        *smax_io = smax;
        int nconv = 0;
        for (int i = 0; i < nu; ++i) {
            const double ratio = fabs(sig_prev[i] - sig[i]) / sig[i];
            if (fabs(res[i]) < tol * smax && ratio < svtol) ++nconv;
        }
        // invariant subspace: |F| at rounding level => every Ritz value is exact, nothing left to iterate on
        // (work clamped to min(m, n) spans the whole space in the first sweep; the lineage code would restart on noise)
        const int converged = (nconv >= nu) || (last_diag == 0.0) || (RF <= 1000.0 * 2.220446049250313e-16 * smax);
        int k = k_in;
        if (k < nu + nconv) k = nu + nconv;
        if (k > w - 3) k = w - 3;
        if (k < 1) k = 1;
        sh_k = k;
        sh_conv = converged;
        status->converged = converged;
        status->nconv = nconv;
        status->k = k;
        status->sweeps = sweeps;
        status->sigma0 = sig[0];
        status->RF = RF;
    }
    __syncthreads();
    const int k = sh_k;
    if (!sh_conv) {
        // next projected matrix: B = 0 ; B_ii = sigma_i ; B_{i,k} = res_i (i < k)
        for (int idx = tid; idx < w * w; idx += nthreads) {
            const int c = idx / w, i = idx - c * w;
            double v = 0.0;
            if (i < k && c == i) v = sig[i];
            if (i < k && c == k) v = res[i];
            B[idx] = v;
        }
        for (int i = tid; i < w; i += nthreads) sig_prev[i] = sig[i];
    }
}

size_t bsvd_smem_bytes(int w) { return ((size_t)2 * w * (w | 1) + 2 * (size_t)w) * sizeof(double); }

bool bsvd_supported(int w) { return w <= 256 && bsvd_smem_bytes(w) + 2048 <= ctx().smem_optin; }

void bsvd_launch(int w, int nu, double *B, double *P, double *Q, double *sig, double *sig_prev, const double *nrm2F,
                 double *smax_io, double tol, double svtol, int k_in, const int *flag_dev, BsvdStatus *status_dev) {
    const size_t smem = bsvd_smem_bytes(w);
    const int pairs = (w + 1) / 2;
    static const bool wide = !(getenv("SVB_BSVD_LP") && atoi(getenv("SVB_BSVD_LP")) == 16);
    KTimer kt(SVB_K_VECTOR, 8.0 * 3 * w * w);
    if (pairs <= 32 && wide) {  // a warp per pair
        if (smem > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(bsvd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int warps = std::max(4, pairs);
        bsvd_kernel<32><<<1, warps * 32, smem, ctx().stream>>>(w, nu, B, P, Q, sig, sig_prev, nrm2F, smax_io, tol, svtol, k_in, flag_dev, status_dev);
    } else {                    // a half-warp per pair: ceil(pairs/2) warps, at most 32
        if (smem > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(bsvd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int warps = std::min(32, std::max(4, (pairs + 1) / 2));
        bsvd_kernel<16><<<1, warps * 32, smem, ctx().stream>>>(w, nu, B, P, Q, sig, sig_prev, nrm2F, smax_io, tol, svtol, k_in, flag_dev, status_dev);
    }
    SVB_LAUNCH_CHECK();
}

}  // namespace svb

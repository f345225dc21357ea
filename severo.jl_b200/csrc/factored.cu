// factored.cu — the scaled HVG operator over the RAW COUNTS ("the scaled matrix never materialised").
//
// The matrix the reference feeds to IRLBA is, entry by entry (normalize.jl:27,36 ; scaling.jl:211-212),
//     B_ij = min( log1p(sf * c_ij / s_i) / sd_j ,  scale_max + mean_j / sd_j )
// with c_ij the integer count, s_i the library size of cell i and (mean_j, sd_j) the moments of gene j. It
// factors as  t_i[c] * (1/sd_j)  with  t_i[c] = log1p(sf*c/s_i): a per-cell table over the few distinct counts
// times a per-gene scale. The explicit layouts of operator.cu stream 10 bytes per nonzero (f64 value + u16 index);
// here a nonzero is ONE 16-bit code and the value is rebuilt from two tables that live in shared memory:
//
//   forward  y_i = sum_l t_i[l] * ( sum_{j in G_il} x_j/sd_j ) - mu.x        G_il = genes of cell i with count l
//            layout: CSR by cell, the genes of a row grouped by count level, every group padded to whole 8-code
//            chunks (pad code = n, xs[n] = 0); per row L group ends (u16, chunk units) and L table entries.
//   adjoint  (S'w)_j = (1/sd_j) * sum_i T[i,c_ij] - (sum w) mu_j             T[i,l] = t_i[l] * w_i
//            layout: cells tiled by R (R*L = 8192 table entries = 64 KB of shared memory), gene-major inside a
//            tile, code = (c-1)*R + i_local, (tile, gene) segments padded to whole chunks (pad code = R*L -> 0).
//
// Entries that do not fit (count > L, count <= 0, or clipped by scale_max) are rare; each becomes one "exception
// chunk" of the same streams that carries its exact scaled value as a Float64 (16 B instead of 2 B for that entry).
// Values: t*(1/sd) instead of t/sd, i.e. every entry within 2 ulp of the reference's (documented in DESIGN.md; far
// inside the 1e-6 bar on sigma).
#include "svb_internal.h"
#include "layout.cuh"
#include "assign.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>

using namespace svb;

svb_factored_s::~svb_factored_s() {
    void *ptrs[] = {tlev, tlevA, inv, f_rowptr, f_code, f_meta, a_gptr, a_code, a_meta, a_slices, counters, partial,
                    fwd_ranges, fwd_rows, e_segptr, e_segsum};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    delete exc;
    delete excT;
}

namespace svb {

constexpr int FCH = 16;   // codes per chunk (one 32-byte load: LDG.256); round 1-2a: 8
constexpr int FEXC = 63;  // level field of an exception chunk in the forward stream

static inline unsigned fgrid(int64_t n, int threads = 256, int max_blocks = 148 * 16) {
    int64_t b = (n + threads - 1) / threads;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(b, max_blocks));
}

// ---------------------------------------------------------------------------------------------
// build kernels
// ---------------------------------------------------------------------------------------------
// per gene: sd, stored mu = mean/sd, clip = scale_max + mu (scaling.jl:205-212, same arithmetic as svb_scale), 1/sd
__global__ void fact_prepare_kernel(const double *__restrict__ mean, const double *__restrict__ var, int64_t n, double scale_max,
                                    double *__restrict__ sd, double *__restrict__ mus, double *__restrict__ cap,
                                    double *__restrict__ inv, int *__restrict__ bad) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double s = sqrt(var[j]);
    if (!(s > 0.0) || !isfinite(s)) *bad = 1;
    sd[j] = s;
    const double m1 = __ddiv_rn(mean[j], s);
    mus[j] = m1;
    cap[j] = scale_max + m1;
    inv[j] = __ddiv_rn(1.0, s);
}

// t_i[l] = log1p((sf * (l+1)) / s_i): the arithmetic of libnorm_kernel<double> (normalize.jl:27,36)
// written twice: cell-major tlev[i*L + l] (forward: one row's table is contiguous) and, per adjoint tile, level-major
// tlevA[tile*R*L + l*R + i_local] (the shared-memory table of the adjoint is level-major so that the bank of an entry is
// its CELL, not its level: 63 % of the entries have count 1 and would otherwise all hit the same bank)
__global__ void fact_tlev_kernel(const long long *__restrict__ s, int64_t m, int log2L, int log2R, int64_t rc, double sf,
                                 double *__restrict__ tlev, double *__restrict__ tlevA) {
    const int64_t total = m << log2L;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = k >> log2L;
        const int l = (int)(k & ((1 << log2L) - 1));
        const long long si = s[i];
        double v = 0.0;
        if (si > 0) v = log1p(__ddiv_rn(__dmul_rn(sf, (double)(l + 1)), (double)si));
        tlev[k] = v;
        const int64_t t = i / rc, il = i - t * rc;  // rc cells per tile, in a table with room for 1 << log2R
        tlevA[(t << (log2R + log2L)) + ((int64_t)l << log2R) + il] = v;
    }
}

__device__ __forceinline__ double fact_block_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += red[i];  // fixed order
    return t;
}

__device__ __forceinline__ double fact_value(int c, int32_t r, int log2L, const double *__restrict__ tlev,
                                             const long long *__restrict__ libsize, double sf) {
    if (c >= 1 && c <= (1 << log2L)) return __ldg(tlev + (((int64_t)r) << log2L) + (c - 1));
    return log1p(__ddiv_rn(__dmul_rn(sf, (double)c), (double)libsize[r]));
}

// Parallel two-pass moments of the log-normalised columns straight from the counts (used when the caller does not
// supply the order-exact Welford moments of svb_mean_var: a fused PCA call keeps the moments internal, and the
// sequential chain of the densest gene alone costs ~150 ms at 1.3 M cells). One CTA per gene, fixed-order sums.
// pass 1: sum[j] = sum of the stored values (this rank's cells); pass 2: ss[j] = sum (v-mean)^2 + (#zeros)*mean^2.
__global__ void __launch_bounds__(512) fact_moments_sum_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                               const int32_t *__restrict__ val, int log2L,
                                                               const double *__restrict__ tlev, const long long *__restrict__ libsize,
                                                               double sf, double *__restrict__ sum) {
    __shared__ double red[32];
    const int64_t j = blockIdx.x;
    const int64_t beg = colptr[j], end = colptr[j + 1];
    double p = 0.0;
    for (int64_t k = beg + threadIdx.x; k < end; k += blockDim.x) p += fact_value(val[k], rowidx[k], log2L, tlev, libsize, sf);
    const double t = fact_block_sum(p, red);
    if (threadIdx.x == 0) sum[j] = t;
}

__global__ void __launch_bounds__(512) fact_moments_ss_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                              const int32_t *__restrict__ val, int log2L,
                                                              const double *__restrict__ tlev, const long long *__restrict__ libsize,
                                                              double sf, int64_t nrow_local, const double *__restrict__ sum,
                                                              const double *__restrict__ mtot, double *__restrict__ mean,
                                                              double *__restrict__ ss) {
    __shared__ double red[32];
    const int64_t j = blockIdx.x;
    const int64_t beg = colptr[j], end = colptr[j + 1];
    const double mu = __ddiv_rn(sum[j], *mtot);
    double q = 0.0;
    for (int64_t k = beg + threadIdx.x; k < end; k += blockDim.x) {
        const double d = fact_value(val[k], rowidx[k], log2L, tlev, libsize, sf) - mu;
        q = fma(d, d, q);
    }
    const double t = fact_block_sum(q, red);
    if (threadIdx.x == 0) {
        mean[j] = mu;
        ss[j] = fma((double)(nrow_local - (end - beg)), mu * mu, t);
    }
}

__global__ void fact_moments_var_kernel(const double *__restrict__ ss, const double *__restrict__ mtot, int64_t n,
                                        double *__restrict__ var) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) var[j] = __ddiv_rn(ss[j], *mtot - 1.0);
}

// how many counts exceed 4 / 8 / 16 / 32 (chooses L)
__global__ void fact_hist_kernel(const int32_t *__restrict__ val, int64_t nnz, unsigned long long *__restrict__ out) {
    unsigned long long c4 = 0, c8 = 0, c16 = 0, c32 = 0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        const int c = val[k];
        c4 += (c > 4 || c < 1);
        c8 += (c > 8 || c < 1);
        c16 += (c > 16 || c < 1);
        c32 += (c > 32 || c < 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c4 += __shfl_xor_sync(0xffffffffu, c4, o);
        c8 += __shfl_xor_sync(0xffffffffu, c8, o);
        c16 += __shfl_xor_sync(0xffffffffu, c16, o);
        c32 += __shfl_xor_sync(0xffffffffu, c32, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out + 0, c4);
        atomicAdd(out + 1, c8);
        atomicAdd(out + 2, c16);
        atomicAdd(out + 3, c32);
    }
}

// level of every stored entry, CSC order: 1..L = count level, 0 = exception (kept with its exact value).
// grid (ncol, FSL/8), 8 warps per block: a gene's column is cut into FSL contiguous slices, one warp each, and the
// number of exceptions of every slice is recorded so that the exception matrix can be filled in order, in parallel.
constexpr int FSL = 64;
__global__ void __launch_bounds__(256) fact_classify_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                            const int32_t *__restrict__ val, int log2L,
                                                            const double *__restrict__ tlev, const double *__restrict__ sd,
                                                            const double *__restrict__ cap, uint8_t *__restrict__ lvl,
                                                            int64_t *__restrict__ excnt) {
    const int64_t j = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int sl = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int64_t beg = colptr[j], len = colptr[j + 1] - beg;
    const int64_t k0 = beg + (len * sl) / FSL, k1 = beg + (len * (sl + 1)) / FSL;
    const double s = sd[j], cp = cap[j];
    const int L = 1 << log2L;
    int nexc = 0;
    for (int64_t k = k0 + lane; k < k1; k += 32) {
        const int c = val[k];
        int lv = 0;
        if (c >= 1 && c <= L) {
            const double t = __ldg(tlev + (((int64_t)rowidx[k]) << log2L) + (c - 1));
            const double v = __ddiv_rn(t, s);  // scaling.jl:211
            lv = (v > cp) ? 0 : c;             // clipped entries keep their exact (clipped) value as exceptions
        }
        lvl[k] = (uint8_t)lv;
        nexc += (lv == 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nexc += __shfl_xor_sync(0xffffffffu, nexc, o);
    if (lane == 0) excnt[j * FSL + sl] = nexc;
}

// the exception matrix (CSC, cells ascending inside a gene): every warp compacts the level-0 entries of its slice,
// in order, at the offset the scan of the slice counts gives it; values are exact and pre-multiplied by sd
__global__ void __launch_bounds__(256) fact_exception_fill_kernel(const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                                              const int32_t *__restrict__ val, const uint8_t *__restrict__ lvl,
                                                              const long long *__restrict__ libsize, double sf,
                                                              const double *__restrict__ sd, const double *__restrict__ cap,
                                                              const int64_t *__restrict__ exoff, int32_t *__restrict__ erow,
                                                              double *__restrict__ eval) {
    const int64_t j = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int sl = blockIdx.y * 8 + (threadIdx.x >> 5);
    int64_t pos = exoff[j * FSL + sl];
    if (exoff[j * FSL + sl + 1] == pos) return;
    const int64_t beg = colptr[j], len = colptr[j + 1] - beg;
    const int64_t k0 = beg + (len * sl) / FSL, k1 = beg + (len * (sl + 1)) / FSL;
    const double s = sd[j], cp = cap[j];
    for (int64_t kb = k0; kb < k1; kb += 32) {
        const int64_t k = kb + lane;
        const bool ex = (k < k1) && (lvl[k] == 0);
        const unsigned bal = __ballot_sync(0xffffffffu, ex);
        if (ex) {
            const int32_t r = rowidx[k];
            const double t = __dmul_rn(sf, (double)val[k]);
            const double v = log1p(__ddiv_rn(t, (double)libsize[r]));
            const double y = __ddiv_rn(v, s);
            const int64_t dst = pos + __popc(bal & ((1u << lane) - 1u));
            erow[dst] = r;
            eval[dst] = ((y > cp) ? cp : y) * s;  // stored pre-multiplied by sd: the kernels apply 1/sd to everything
        }
        pos += __popc(bal);
    }
}

// ---- bank-aware placement ---------------------------------------------------------------------------------------
// The product kernels gather 8-byte table entries from shared memory; a half-warp's 16 gathers are served in one pass
// only if they fall in 16 different 8-byte banks (bank = index mod 16), and with the entries in ascending order the
// measured cost is ~3 passes per 16 gathers (ncu: 2/3 of the shared-load wavefronts are bank conflicts), which is what
// bounds the kernels. The ORDER of the entries inside a (row, level) group / (tile, gene) segment is free, so the
// builders choose it: the 16 lanes of a half-warp read element e of 16 consecutive chunks (aligned to 16 in the global
// chunk index: the kernels start every warp on a multiple of 32 chunks) -- a "set". Entries are dealt to the sets in
// round-robin order over the residues (the j-th entry of every residue class, class after class), and the slots are
// filled set by set: any 16 consecutive entries of that sequence have different residues as long as no class has run
// out, so the conflicts are confined to the tail of a group. Deterministic (ranks follow the ascending order).
// position of the j-th entry (0-based) of residue class rho in the round-robin sequence; cnt[16] = class sizes
__device__ __forceinline__ int rr_position(const int *__restrict__ cnt, int rho, int j) {
    int p = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int c = cnt[q];
        p += min(c, j) + ((q < rho && c > j) ? 1 : 0);
    }
    return p;
}
// slot (index into the u16 code array) of sequence position p for a group that owns the global chunks [g0, g1):
// the slots are enumerated set by set -- aligned block of 16 chunks, then element e, then chunk
__device__ __forceinline__ int64_t rr_slot(int64_t g0, int64_t g1, int p) {
    const int w0 = (int)(min(g1, ((g0 >> 4) + 1) << 4) - g0);  // chunks of the group in its first block (1..16)
    if (p < FCH * w0) return (g0 + p % w0) * FCH + p / w0;
    const int pp = p - FCH * w0;
    const int64_t cb = g0 + w0 + ((pp / (16 * FCH)) << 4);      // every block before the last is full (16 chunks x FCH slots)
    const int wb = (int)min((int64_t)16, g1 - cb);
    const int q = pp % (16 * FCH);
    return (cb + q % wb) * FCH + q / wb;
}

// ---- replica assignment (round 2) --------------------------------------------------------------------------------
// What the round-robin order leaves (the tails of the groups, sets shared by two groups / segments: 1.9 passes per set in the
// forward stream, 1.55 in the adjoint at C3) is removed with bank-shifted REPLICAS of the gathered table: replica r of a table
// region starts 5r banks later, so an entry can be read from any of nrep banks, and the builder decides which — per SET (element
// e of an aligned block of 16 chunks = the 16 gathers a half-warp issues together), one thread per set: entries with a single
// copy take their bank first; the others are matched to the free banks by augmenting paths (Kuhn's bipartite matching, 16 x 16);
// what cannot be matched goes to its least-loaded candidate. The chosen replica is written into the code, so the product
// kernels are unchanged apart from a larger table. Host simulation on the C3 matrix (tools/studies/bank_sim.py): forward,
// 4 replicas of xs (64 KB): 1.91 -> 1.10 passes per set; adjoint, 4 replicas of levels 1-2 only (176 KB table): 1.55 -> 1.10.
// The pads of a set share one address and count as ONE entry. Exception chunks are left untouched.
// Two kernels: the FAST one (greedy, registers only) handles most sets and appends the rest to a list; the FULL one runs the
// augmenting-path matching (local arrays) on that list only. (One kernel doing both kept its 16-entry arrays in local memory
// for every set: 55 + 75 ms of a 195 ms operator build at C3.)
template <bool FULL, int EFFORT>
__global__ void __launch_bounds__(256) fact_assign_kernel(uint16_t *__restrict__ code, const uint8_t *__restrict__ meta,
                                                          int64_t nchunks, AssignGeom G, unsigned long long *__restrict__ stats,
                                                          unsigned int *__restrict__ hard_list, unsigned int hard_n, bool commit_all) {
    const int64_t nsets = FULL ? (int64_t)hard_n : ((nchunks + 15) >> 4) * FCH;
    unsigned long long mypasses = 0, mysets = 0;
    for (int64_t si = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; si < nsets; si += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = FULL ? (int64_t)hard_list[si] : si;
        const int64_t blk = s / FCH;
        const int e = (int)(s % FCH);
        int v[16];
        unsigned present = 0;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            v[c] = 0;
            const int64_t ch = (blk << 4) + c;
            if (ch < nchunks) {
                v[c] = code[ch * FCH + e];
                present |= 1u << c;
            }
        }
        if (!present) continue;
        int p;
        if (FULL) {
            p = assign_set_full(G, v, present);
        } else {
            p = assign_set_fast(G, v, present, commit_all, EFFORT);
            if (p < 0) {  // one atomic per warp (the counter lives behind the list)
                const unsigned act = __activemask();
                const int leader = __ffs((int)act) - 1, lane = threadIdx.x & 31;
                unsigned int at = 0;
                if (lane == leader) at = atomicAdd(&hard_list[nsets], (unsigned int)__popc(act));
                at = __shfl_sync(act, at, leader);
                hard_list[at + __popc(act & ((1u << lane) - 1u))] = (unsigned int)s;
                continue;
            }
        }
        mypasses += (unsigned long long)p;
        mysets += 1;
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if ((present >> c) & 1u) code[((blk << 4) + c) * FCH + e] = (uint16_t)v[c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mypasses += __shfl_xor_sync(0xffffffffu, mypasses, o);
        mysets += __shfl_xor_sync(0xffffffffu, mysets, o);
    }
    if ((threadIdx.x & 31) == 0 && mysets) {
        atomicAdd(stats + 0, mypasses);
        atomicAdd(stats + 1, mysets);
    }
}

// Default: the sets the greedy pass leaves with a bank conflict (about half of them at C3) go through the augmenting-path
// matching (C3: 4 + 10 ms per stream, passes per set 1.49 / 1.55 -> 1.10 / 1.17, 9.6 ms off one IRLBA solve).
// SVB_FACT_MATCH=greedy0 | greedy1 | greedy keeps the greedy pass alone (its steps up to 3 / 2+3 / all; assign.cuh) for an
// operator that is used for a handful of products only.
static bool fact_full_matching() {
    static const bool full = !getenv("SVB_FACT_MATCH") || !strcmp(getenv("SVB_FACT_MATCH"), "full");
    return full;
}

static int fact_greedy_effort() {  // SVB_FACT_MATCH=greedy0 / greedy1: the greedy pass without its steps 4 / 2+4 (assign.cuh)
    static const int e = !getenv("SVB_FACT_MATCH") ? 2 : !strcmp(getenv("SVB_FACT_MATCH"), "greedy0") ? 0 : !strcmp(getenv("SVB_FACT_MATCH"), "greedy1") ? 1 : 2;
    return e;
}

// rewrites the stream with the chosen replicas; returns the average passes per set
static double run_assign(uint16_t *code, const uint8_t *meta, int64_t nchunks, const AssignGeom &G, cudaStream_t st) {
    if (nchunks <= 0) return 0.0;
    const bool full = fact_full_matching();
    const int64_t nsets = ((nchunks + 15) >> 4) * FCH;
    SVB_CHECK(nsets < 4000000000ll, SVB_EDIM, "count-level operator: stream too long for the replica assignment");
    DevBuf<unsigned long long> d(2);
    DevBuf<unsigned int> hard(full ? (size_t)nsets + 1 : 1);
    SVB_CUDA(cudaMemsetAsync(d.p, 0, 2 * sizeof(unsigned long long), st));
    if (full) SVB_CUDA(cudaMemsetAsync(hard.p + nsets, 0, sizeof(unsigned int), st));
    const bool timing = getenv("SVB_FACT_TIMING") != nullptr;
    auto now = [&]() { cudaStreamSynchronize(st); return std::chrono::steady_clock::now(); };
    auto t0 = timing ? now() : std::chrono::steady_clock::time_point();
    const int effort = full ? 0 : fact_greedy_effort();
    const int g1 = fgrid(nsets, 256, 148 * 32);
    if (effort == 0) fact_assign_kernel<false, 0><<<g1, 256, 0, st>>>(code, meta, nchunks, G, d.p, hard.p, 0u, !full);
    else if (effort == 1) fact_assign_kernel<false, 1><<<g1, 256, 0, st>>>(code, meta, nchunks, G, d.p, hard.p, 0u, !full);
    else fact_assign_kernel<false, 2><<<g1, 256, 0, st>>>(code, meta, nchunks, G, d.p, hard.p, 0u, !full);
    unsigned int nhard = 0;
    if (full) {
        SVB_CUDA(cudaMemcpyAsync(&nhard, hard.p + nsets, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
    }
    auto t1 = timing ? now() : t0;
    if (nhard) fact_assign_kernel<true, 0><<<fgrid(nhard, 256, 148 * 32), 256, 0, st>>>(code, meta, nchunks, G, d.p, hard.p, nhard, false);
    if (timing) {
        auto t2 = now();
        fprintf(stderr, "[svb counts build]     assignment: %lld sets, greedy %.3f ms, %u sets to the matching %.3f ms\n", (long long)nsets,
                std::chrono::duration<double, std::milli>(t1 - t0).count(), nhard, std::chrono::duration<double, std::milli>(t2 - t1).count());
    }
    count_launch(2);
    SVB_LAUNCH_CHECK();
    unsigned long long h[2];
    SVB_CUDA(cudaMemcpyAsync(h, d.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    return h[1] ? (double)h[0] / (double)h[1] : 0.0;
}

// adjoint layout, pass 1: chunks per (tile, gene) segment (at least one: the stream kernels count segments by their
// "last chunk" flags, so an empty segment is one all-pad chunk). Sub-warps of 8 lanes, one segment each.
__global__ void __launch_bounds__(256) fact_seg_count_kernel(const int64_t *__restrict__ startpos, const uint8_t *__restrict__ lvl,
                                                             const int64_t *__restrict__ estart, int64_t nseg, int64_t ncol,
                                                             int64_t *__restrict__ gptr) {
    const int lane = threadIdx.x & 31, sl = lane & 7;
    const unsigned submask = 0xffu << (lane & ~7);
    const int64_t sub = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int64_t nsub = ((int64_t)gridDim.x * blockDim.x) >> 3;
    for (int64_t s = sub; s < nseg; s += nsub) {
        const int64_t b = startpos[s], e = startpos[s + ncol];
        int c = 0;
        for (int64_t k = b + sl; k < e; k += 8) c += (lvl[k] != 0);
        c += __shfl_xor_sync(submask, c, 4);
        c += __shfl_xor_sync(submask, c, 2);
        c += __shfl_xor_sync(submask, c, 1);
        (void)estart;
        if (sl == 0) gptr[s] = max(1, (c + FCH - 1) / FCH);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) gptr[nseg] = 0;
}

// adjoint layout, pass 2: stable placement (cells ascending) of code = (level-1)*R + i_local, pads at the segment end,
// and the per-chunk flag "last chunk of its segment"
__global__ void __launch_bounds__(256) fact_seg_fill_kernel(const int64_t *__restrict__ startpos, const int32_t *__restrict__ rowidx,
                                                            const uint8_t *__restrict__ lvl, int64_t nseg, int64_t ncol,
                                                            int log2R, int64_t rc, int log2L, const int64_t *__restrict__ gptr,
                                                            const int64_t *__restrict__ estart, const int32_t *__restrict__ erow,
                                                            const double *__restrict__ eval, uint16_t *__restrict__ code,
                                                            uint8_t *__restrict__ meta) {
    __shared__ int scnt[32][16], srun[32][16];  // per sub-warp: size / running rank of every residue class (cell mod 16)
    const int lane = threadIdx.x & 31, sl = lane & 7;
    const int shift = lane & ~7;
    const unsigned submask = 0xffu << shift;
    const int sw = threadIdx.x >> 3;
    int *cnt = scnt[sw], *run = srun[sw];
    const int64_t sub = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int64_t nsub = ((int64_t)gridDim.x * blockDim.x) >> 3;
    const uint16_t padcode = (uint16_t)(1u << (log2R + log2L));
    // every sub-warp runs the same number of outer iterations (the loop bound does not depend on the lane)
    for (int64_t s = sub; s < nseg; s += nsub) {
        const int64_t b = startpos[s], e = startpos[s + ncol];
        const int64_t t = s / ncol;
        const int64_t row0 = t * rc;
        const int64_t c0 = gptr[s], c1 = gptr[s + 1];
        (void)estart; (void)erow; (void)eval;
        const int64_t cx = c1;
        cnt[sl] = 0; cnt[sl + 8] = 0; run[sl] = 0; run[sl + 8] = 0;
        for (int64_t p = c0 * FCH + sl; p < cx * FCH; p += 8) code[p] = padcode;
        __syncwarp(submask);
        for (int64_t k = b + sl; k < e; k += 8)
            if (lvl[k]) atomicAdd(&cnt[(rowidx[k] - row0) & 15], 1);
        __syncwarp(submask);
        for (int64_t k0 = b; k0 < e; k0 += 8) {
            const int64_t k = k0 + sl;
            int lv = 0, il = 0;
            if (k < e) {
                lv = lvl[k];
                il = (int)(rowidx[k] - row0);
            }
            const int rho = lv ? (il & 15) : (16 + sl);  // lanes without an entry match nobody
            const unsigned peers = (__match_any_sync(submask, rho) >> shift) & 0xffu;
            if (lv) {
                const int j = run[rho] + __popc(peers & ((1u << sl) - 1u));
                code[rr_slot(c0, cx, rr_position(cnt, rho, j))] = (uint16_t)(((lv - 1) << log2R) + il);
            }
            __syncwarp(submask);
            if (lv && (peers >> sl) == 1u) run[rho] += __popc(peers);  // the last lane of every class updates its counter
            __syncwarp(submask);
        }
        for (int64_t c = c0 + sl; c < c1; c += 8) meta[c] = (uint8_t)(c == c1 - 1);
        __syncwarp(submask);
    }
}

// adjoint: gene boundaries that cut every tile into K slices (one per warp of the CTA) of equal WORK. A warp iteration streams
// 32 chunks in ~120 instructions, the end of a (tile, gene) segment costs ~50 more (reduction across the lanes + store), i.e.
// as much as ~12 chunks: slices of equal chunk count made the warps that own the sparse genes (many short segments) arrive
// late at the tile barrier (9 % of the kernel's stall samples, profiles/r04_spmv.md). Work(g) = chunks before gene g + alpha * g.
__global__ void fact_slices_kernel(const int64_t *__restrict__ gptr, int64_t ntiles, int64_t n, int K, int alpha, int32_t *__restrict__ slices) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntiles * (K + 1)) return;
    const int64_t t = i / (K + 1);
    const int k = (int)(i - t * (K + 1));
    const int64_t *gp = gptr + t * n;
    const int64_t c0 = gp[0], c1 = gp[n];
    int64_t g;
    if (k == 0) g = 0;
    else if (k == K) g = n;
    else {
        const int64_t total = (c1 - c0) + (int64_t)alpha * n;
        const int64_t target = (int64_t)(((__int128)total * k) / K);
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((gp[mid] - c0) + (int64_t)alpha * mid < target) lo = mid + 1; else hi = mid;
        }
        g = lo;
    }
    slices[i] = (int32_t)g;
}

// forward layout, pass 1 (CSR by cell with the level of every entry): per row the cumulative group ends in
// chunks (u16) and the row's chunk count (at least one chunk per row, see above). One warp per row.
__global__ void __launch_bounds__(256) fact_row_hist_kernel(const int64_t *__restrict__ rowptr, const uint8_t *__restrict__ rlvl,
                                                            int64_t m, int L, uint16_t *__restrict__ gend,
                                                            int64_t *__restrict__ rowchunks) {
    __shared__ int hist[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * 8 + w;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    for (int64_t r = warp; r < m; r += nwarps) {
        hist[w][lane] = 0;
        __syncwarp();
        const int64_t b = rowptr[r], e = rowptr[r + 1];
        int ne = 0;  // exceptions of the row (level byte 0): one chunk each, after the level groups
        for (int64_t k = b + lane; k < e; k += 32) {
            const int lv = rlvl[k];
            if (lv) atomicAdd(&hist[w][lv - 1], 1);
            else ++ne;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ne += __shfl_xor_sync(0xffffffffu, ne, o);
        __syncwarp();
        int ch = (lane < L) ? (hist[w][lane] + FCH - 1) / FCH : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, ch, o);
            if (lane >= o) ch += y;
        }
        const int total = __shfl_sync(0xffffffffu, ch, 31);
        (void)ne;                     // exception entries live in the side matrix, not in the stream
        if (total == 0) ch = 1;       // row without coded entries: one all-pad chunk in level 0 (it carries the row end)
        if (lane < L) gend[r * L + lane] = (uint16_t)ch;
        if (lane == 31) rowchunks[r] = ch;
        __syncwarp();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) rowchunks[m] = 0;
}

// forward layout, pass 2: the genes of every (row, level) group in the bank-aware order (see above), pads (gene index n,
// xs[n] = 0) in the unused slots, and the per-chunk byte (level << 1) | last-chunk-of-the-row. One warp per row;
// dynamic shared memory: per warp L x 16 class sizes and L x 16 running ranks.
__global__ void __launch_bounds__(256) fact_row_place_kernel(const int64_t *__restrict__ rowptr, const uint16_t *__restrict__ ridx,
                                                             const uint8_t *__restrict__ rlvl, int64_t m, int L, int padgene,
                                                             int cshift, const int64_t *__restrict__ frowptr,
                                                             const uint16_t *__restrict__ gend, const int64_t *__restrict__ ecolptr,
                                                             const int32_t *__restrict__ erow, const double *__restrict__ eval,
                                                             uint16_t *__restrict__ code, uint8_t *__restrict__ meta) {
    extern __shared__ int smi[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int *cnt = smi + (size_t)wid * 2 * L * 16;  // [L][16]
    int *run = cnt + L * 16;                      // [L][16]
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < m; r += nwarps) {
        const int64_t b = rowptr[r], e = rowptr[r + 1];
        const int64_t cbase = frowptr[r];
        const int total = (int)(frowptr[r + 1] - cbase);
        const int coded = gend[r * L + L - 1];  // chunks of the level groups (the exception chunks follow)
        for (int i = lane; i < 2 * L * 16; i += 32) cnt[i] = 0;
        for (int64_t p = cbase * FCH + lane; p < (cbase + coded) * FCH; p += 32) code[p] = (uint16_t)(padgene << cshift);
        {
            int prev = 0;
            for (int l = 0; l < L; ++l) {
                const int ge = gend[r * L + l];
                for (int c = prev + lane; c < ge; c += 32) meta[cbase + c] = (uint8_t)((l << 1) | (c == total - 1));
                prev = ge;
            }
        }
        __syncwarp();
        for (int64_t k = b + lane; k < e; k += 32) {
            const int lv = rlvl[k];
            if (lv) atomicAdd(&cnt[(lv - 1) * 16 + (ridx[k] & 15)], 1);
        }
        __syncwarp();
        for (int64_t k0 = b; k0 < e; k0 += 32) {
            const int64_t k = k0 + lane;
            int lv = 0, g = 0;
            if (k < e) {
                lv = rlvl[k];
                g = ridx[k];
            }
            const int key = lv ? ((lv - 1) * 16 + (g & 15)) : (L * 16 + lane);  // lanes without a coded entry match nobody
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            if (lv) {
                const int l = lv - 1;
                const int j = run[key] + __popc(peers & ((1u << lane) - 1u));
                const int64_t g0 = cbase + (l ? gend[r * L + l - 1] : 0), g1 = cbase + gend[r * L + l];
                code[rr_slot(g0, g1, rr_position(cnt + l * 16, g & 15, j))] = (uint16_t)(g << cshift);  // index or byte offset
            }
            __syncwarp();
            if (lv && (peers >> lane) == 1u) run[key] += __popc(peers);  // the last lane of every class updates its counter
            __syncwarp();
        }
        (void)ecolptr; (void)erow; (void)eval;
        __syncwarp();
    }
}

// warp b of NW streams the chunks of the rows [wrow[b], wrow[b+1]): whole rows, equal chunk counts
__global__ void fact_warp_ranges_kernel(const int64_t *__restrict__ frowptr, int64_t m, int64_t nchunks, int NW,
                                        int64_t *__restrict__ wstart, int64_t *__restrict__ wrow) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > NW) return;
    int64_t r;
    if (b == 0) r = 0;
    else if (b == NW) r = m;
    else {
        const int64_t target = (int64_t)(((__int128)nchunks * b) / NW);
        int64_t lo = 0, hi = m;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (frowptr[mid] < target) lo = mid + 1; else hi = mid;
        }
        r = lo;
    }
    wrow[b] = r;
    wstart[b] = frowptr[r];
}

// ---------------------------------------------------------------------------------------------
// products: chunk-stream kernels. A warp streams a contiguous range of 16-byte chunks (32 chunks = 512 B per
// iteration, loads issued PD iterations ahead), every lane turns its chunk into one partial value, and a segmented
// warp scan over the "last chunk" flags adds the partials of a row / (tile, gene) segment; a segment that continues
// into the next iteration is carried in a register. No per-row setup, no idle lanes on short rows, fixed summation
// order (deterministic), and the bytes in flight do not depend on the row lengths.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double fblock_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

__device__ __forceinline__ double gather8(const double *__restrict__ T, const uint4 q) {
    const double a0 = T[q.x & 0xffffu], a1 = T[q.x >> 16], a2 = T[q.y & 0xffffu], a3 = T[q.y >> 16];
    const double a4 = T[q.z & 0xffffu], a5 = T[q.z >> 16], a6 = T[q.w & 0xffffu], a7 = T[q.w >> 16];
    return ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// one chunk = 16 codes = 32 bytes, fetched with ONE 256-bit load (LDG.E.256 on sm_100)
struct __align__(32) Chunk {
    uint4 lo, hi;
};
__device__ __forceinline__ Chunk ld_chunk(const Chunk *p) {  // streamed once: do not keep it in L1
    Chunk c;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(c.lo.x), "=r"(c.lo.y), "=r"(c.lo.z), "=r"(c.lo.w), "=r"(c.hi.x), "=r"(c.hi.y), "=r"(c.hi.z), "=r"(c.hi.w)
                 : "l"(p));
    return c;
}
__device__ __forceinline__ uint4 ld_stream16(const uint4 *p) {  // streamed once: do not keep it in L1
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// inclusive segmented sum over the lanes of a warp; `lastbits` = ballot of the "segment ends at this lane" flags.
// Returns the running sum of the lane's segment up to and including the lane; head0 = the lane's segment starts at lane 0.
__device__ __forceinline__ double seg_scan(double v, unsigned lastbits, int lane, bool &head0) {
    const unsigned heads = (lastbits << 1) | 1u;
    const unsigned le = (lane == 31) ? 0xffffffffu : ((2u << lane) - 1u);
    const int head_lane = 31 - __clz(heads & le);
    const int d = lane - head_lane;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if (d >= o) v += t;
    }
    head0 = (head_lane == 0);
    return v;
}

constexpr int FPD = 3;  // prefetch distance of the stream kernels, in iterations (3 x 16 B per lane in flight)

// gather with codes that are BYTE offsets (index * 8 < 65536): one instruction per code instead of two
__device__ __forceinline__ double gather8b(const double *__restrict__ T, const uint4 q) {
    const char *b = reinterpret_cast<const char *>(T);
#define SVB_AT(off) (*reinterpret_cast<const double *>(b + (off)))
    const double a0 = SVB_AT(q.x & 0xffffu), a1 = SVB_AT(q.x >> 16), a2 = SVB_AT(q.y & 0xffffu), a3 = SVB_AT(q.y >> 16);
    const double a4 = SVB_AT(q.z & 0xffffu), a5 = SVB_AT(q.z >> 16), a6 = SVB_AT(q.w & 0xffffu), a7 = SVB_AT(q.w >> 16);
#undef SVB_AT
    return ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// Exception entries (count > L, count < 1, clipped by scale_max): NOT in the streams. Round 1 carried each as a 16-byte chunk
// of its own — 1.3 % of the entries, but 9 % of the lane slots of a warp iteration and a divergent branch in ~95 % of the
// iterations. Reading them at the row / segment ends inside the stream loop was worse (two dependent global-load latencies in
// the critical path of every end: 0.55 -> 0.71 ms and 0.63 -> 1.48 ms per product). They are two small side matrices instead:
// forward: cell-major, one thread per cell adds alpha * sum(value * x_g / sd_g) to y after the stream kernel (this kernel);
// adjoint: gene-major, summed into the gene's result by the reduce kernel that already adds up the tile partials.
constexpr int EXS = 4096;  // exception entries per segment of the adjoint's side sum
__global__ void exc_segcount_kernel(const int64_t *__restrict__ ecolptr, int64_t n, int64_t *__restrict__ segptr) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) segptr[g] = (ecolptr[g + 1] - ecolptr[g] + EXS - 1) / EXS;
    if (g == n) segptr[g] = 0;
}
// one CTA per segment: segsum[s] = sum over the segment's entries of value * w[cell], fixed order (thread-strided partial sums,
// then the block tree) — deterministic whatever the distribution of the exceptions over the genes
__global__ void __launch_bounds__(256) adj_exceptions_kernel(const int64_t *__restrict__ ecolptr, const int32_t *__restrict__ erow,
                                                             const double *__restrict__ eval, const double *__restrict__ w,
                                                             const int64_t *__restrict__ segptr, int64_t n, double *__restrict__ segsum) {
    __shared__ double red[32];
    const int64_t s = blockIdx.x;
    int64_t lo = 0, hi = n;  // gene of this segment: the last g with segptr[g] <= s
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (segptr[mid] <= s) lo = mid; else hi = mid - 1;
    }
    const int64_t g = lo;
    const int64_t k0 = ecolptr[g] + (s - segptr[g]) * EXS, k1 = min(k0 + EXS, ecolptr[g + 1]);
    double p = 0.0;
    for (int64_t k = k0 + threadIdx.x; k < k1; k += 256) p = fma(__ldg(eval + k), __ldg(w + __ldg(erow + k)), p);
    const double t = fact_block_sum(p, red);
    if (threadIdx.x == 0) segsum[s] = t;
}

// y_i += alpha * sum over the cell's exception entries of value * (x_g / sd_g). A block owns 256 consecutive cells: their
// entries are one contiguous range of the cell-major side matrix, loaded coalesced and turned into products in shared memory
// (2048 at a time); every thread then adds up its own cell's products in order. (First version: a serial loop per thread
// over its cell's entries straight from global memory — 74 us at C3 for 124 MB; uncoalesced.)
constexpr int FXT = 2048;
__global__ void __launch_bounds__(256) fwd_exceptions_kernel(const int64_t *__restrict__ erowptr, const int32_t *__restrict__ egene,
                                                             const double *__restrict__ evalr, int64_t m, const double *__restrict__ x,
                                                             const double *__restrict__ inv, double alpha, double *__restrict__ y) {
    __shared__ double prod[FXT];
    for (int64_t c0 = (int64_t)blockIdx.x * 256; c0 < m; c0 += (int64_t)gridDim.x * 256) {
        const int64_t c1 = min(c0 + 256, m);
        const int64_t ebeg = erowptr[c0], eend = erowptr[c1];
        if (ebeg == eend) continue;  // block-uniform
        const int64_t i = c0 + threadIdx.x;
        const int64_t a = i < m ? erowptr[i] : eend, b = i < m ? erowptr[i + 1] : eend;
        double s = 0.0;
        for (int64_t base = ebeg; base < eend; base += FXT) {
            const int nb = (int)min((int64_t)FXT, eend - base);
            for (int k = threadIdx.x; k < nb; k += 256) {
                const int g = __ldg(egene + base + k);
                prod[k] = __ldg(evalr + base + k) * (__ldg(x + g) * __ldg(inv + g));  // value*sd times x/sd (the table entry xs)
            }
            __syncthreads();
            for (int64_t k = max(a, base), ke = min(b, base + nb); k < ke; ++k) s += prod[k - base];
            __syncthreads();
        }
        if (b > a) y[i] = fma(alpha, s, y[i]);
    }
}

// forward: y_i = alpha*(sum over the row's chunks of t_i[level] * sum_8 xs[gene] - mu.x) + beta*y_i + csign*(*coef)*cvec_i
// BO: the codes are byte offsets (n <= 8190). After the ncu pass that showed the issue slots 70 % busy once the bank
// conflicts were halved, the loop works on 32-bit indices relative to the warp's range and the three pipeline stages
// are unrolled by hand (no register rotation): ~200 -> ~150 instructions per 32 chunks.
template <int BLOCK, bool BO, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
fwd_stream_kernel(const Chunk *__restrict__ code, const uint8_t *__restrict__ meta, const double *__restrict__ tlev, int log2L,
                  const int64_t *__restrict__ wstart, const int64_t *__restrict__ wrow, int64_t n, const double *__restrict__ x,
                  const double *__restrict__ inv, const double *__restrict__ mu, double alpha, double beta, double *__restrict__ y,
                  const double *__restrict__ coef, double csign, const double *__restrict__ cvec, int nrep, int stride) {
    extern __shared__ double smem[];
    double *red = smem;      // 32
    double *xs = smem + 32;  // nrep bank-shifted replicas of {x_j / sd_j, j < n ; 0 (the pad gene)}, `stride` entries apart
    double part = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += BLOCK) {
        const double xv = x[j];
        const double v = xv * inv[j];
        for (int r = 0; r < nrep; ++r) xs[(int64_t)r * stride + j] = v;
        part = fma(mu[j], xv, part);
    }
    if ((int)threadIdx.x < nrep) xs[(int64_t)threadIdx.x * stride + n] = 0.0;
    {
        // the epilogue constants are needed at row ends only: they live in shared memory (red[0..3]) instead of eight
        // registers of the stream loop, which runs at the 64-register limit of two 512-thread CTAs per SM
        const double mudot_ = fblock_sum(part, red);  // its barriers publish xs
        __syncthreads();
        if (threadIdx.x == 0) {
            red[0] = mudot_;
            red[1] = (coef != nullptr) ? csign * (*coef) : 0.0;
            red[2] = alpha;
            red[3] = beta;
        }
        __syncthreads();
    }

    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t gw = (int64_t)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
    const int64_t cbeg = __ldg(wstart + gw), cend = __ldg(wstart + gw + 1);
    // lane = chunk index mod 32: the builder's bank-aware order assumes half-warps read aligned blocks of 16 chunks
    const int64_t cbase = cbeg & ~(int64_t)31;
    const int lo = (int)(cbeg - cbase), hi = (int)(cend - cbase);  // this warp's chunks, relative to cbase
    const int64_t rbase = __ldg(wrow + gw);
    const Chunk *pc = code + cbase + lane;
    const uint8_t *pm = meta + cbase + lane;
    const double *tl = tlev + (rbase << log2L);
    double *yb = y + rbase;
    const double *cvb = cvec ? cvec + rbase : nullptr;
    const unsigned padc = (unsigned)(BO ? n * 8 : n) * 0x10001u;
    Chunk padq;
    padq.lo = make_uint4(padc, padc, padc, padc);
    padq.hi = padq.lo;

    int rel = lane;   // this lane's chunk of the current iteration, relative to cbase
    int rowrel = 0;   // row (relative to rbase) of the first chunk of the next iteration to be decoded
    constexpr int NSF = 2;  // stages in flight: 2 x 32 bytes per lane (round 1: 3 x 16)
    Chunk q0, q1;
    unsigned m0, m1;
    {
        const bool ok0 = rel >= lo && rel < hi, ok1 = rel + 32 < hi;
        q0 = ok0 ? ld_chunk(pc) : padq;
        m0 = ok0 ? (unsigned)__ldg(pm) : 0u;
        q1 = ok1 ? ld_chunk(pc + 32) : padq;
        m1 = ok1 ? (unsigned)__ldg(pm + 32) : 0u;
        pc += 32 * NSF;
        pm += 32 * NSF;
    }
    // decode iteration 0 and fetch its table entries
    unsigned bal0 = __ballot_sync(0xffffffffu, (m0 & 1u) != 0u);
    int row0 = rowrel + __popc(bal0 & lt);
    rowrel += __popc(bal0);
    double t0 = (rel >= lo && rel < hi) ? __ldg(tl + ((int64_t)row0 << log2L) + (m0 >> 1)) : 0.0;
    double acc = 0.0;  // this lane's share of the row that is still open (summed across lanes when the row ends)
    // one iteration: QA/MA = current stage (refilled with iteration +FPD once consumed), MB = meta of the next iteration
#define SVB_FWD_STEP(QA, MA, MB)                                                                                      \
    {                                                                                                                 \
        const unsigned bal1 = __ballot_sync(0xffffffffu, ((MB) & 1u) != 0u);                                          \
        const int row1 = rowrel + __popc(bal1 & lt);                                                                  \
        rowrel += __popc(bal1);                                                                                       \
        const double t1 = (rel + 32 < hi) ? __ldg(tl + ((int64_t)row1 << log2L) + ((MB) >> 1)) : 0.0;                 \
        const double v = t0 * (BO ? (gather8b(xs, (QA).lo) + gather8b(xs, (QA).hi)) : (gather8(xs, (QA).lo) + gather8(xs, (QA).hi))); \
        const unsigned mcur = (MA);                                                                                   \
        if (rel + 32 * NSF < hi) {                                                                                    \
            (QA) = ld_chunk(pc);                                                                                      \
            (MA) = (unsigned)__ldg(pm);                                                                               \
        } else {                                                                                                      \
            (QA) = padq;                                                                                              \
            (MA) = 0u;                                                                                                \
        }                                                                                                             \
        pc += 32;                                                                                                     \
        pm += 32;                                                                                                     \
        const int nends = __popc(bal0); /* rows that end in this iteration (warp-uniform) */                          \
        if (nends == 0) {                                                                                             \
            acc += v; /* the open row continues: lane-private partial sums, no shuffles */                            \
        } else {                                                                                                      \
            double r;                                                                                                 \
            if (nends == 1) { /* the usual case for rows longer than 32 chunks: one butterfly */                      \
                const int b1 = __ffs(bal0) - 1;                                                                       \
                r = acc + ((lane <= b1) ? v : 0.0);                                                                   \
                _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);           \
            } else { /* several (short) rows: carried sums enter at lane 0, then the segmented scan */                \
                double tot = acc;                                                                                     \
                _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);       \
                bool head0;                                                                                           \
                r = seg_scan(v + (lane == 0 ? tot : 0.0), bal0, lane, head0);                                         \
            }                                                                                                         \
            if (mcur & 1u) {                                                                                          \
                r = red[2] * (r - red[0]);                                                                            \
                const double beta_ = red[3];                                                                          \
                if (beta_ != 0.0) r = fma(beta_, yb[row0], r);                                                        \
                if (cvb != nullptr) r = fma(red[1], cvb[row0], r);                                                    \
                yb[row0] = r;                                                                                         \
            }                                                                                                         \
            acc = (lane > 31 - __clz(bal0)) ? v : 0.0; /* the chunks after the last row end open the next row */      \
        }                                                                                                             \
        bal0 = bal1;                                                                                                  \
        row0 = row1;                                                                                                  \
        t0 = t1;                                                                                                      \
        rel += 32;                                                                                                    \
    }
    while (rel - lane < hi) {
        SVB_FWD_STEP(q0, m0, m1)
        if (rel - lane >= hi) break;
        SVB_FWD_STEP(q1, m1, m0)
    }
#undef SVB_FWD_STEP
}

// physical layout of the adjoint tile table (see the replica assignment above); identity = {0, 1, 0, 0, 0, R*L, R*L + 1}
struct AdjGeom {
    int nlr, nrep, strideA, levstride, baseB, pad, wbase;
    int rc;  // cells per tile (<= R = 1 << log2R, a multiple of 16: the bank of an entry stays cell mod 16)
};

// adjoint, stage 1: partial[t][g] = (1/sd_g) * sum over the chunks of segment (t,g) of sum_8 T[code], T[l*R+i] = t_i[l]*w_i;
// partial[t][n] = sum of w over the tile. CTAs take tiles from a global counter (any order gives the same bits: every
// tile has its own partial row); inside a tile warp k streams the k-th slice of the tile's chunks.
// LAZY: lane-private partial sums, cross-lane reduction only in iterations where a segment ends (pays when the segments
// are longer than a warp iteration: the 1024-cell tiles); otherwise a segmented scan every iteration + a carry register.
template <int BLOCK, int MINB, bool LAZY, int NS>
__global__ void __launch_bounds__(BLOCK, MINB)
adj_stream_kernel(const int64_t *__restrict__ gptr, const Chunk *__restrict__ code, const uint8_t *__restrict__ meta,
                  const double *__restrict__ tlevA, int log2L, int log2R, int64_t m, int64_t n, int64_t ntiles,
                  const double *__restrict__ w, const double *__restrict__ inv, double *__restrict__ partial,
                  const int32_t *__restrict__ slices, unsigned int *__restrict__ counters, AdjGeom G) {
    extern __shared__ double smem[];
    __shared__ long long cur_tile, nxt_tile;  // tiles are taken from the counter ONE AHEAD: the next tile's level table is
                                              // prefetched into L2 while this one is streamed (the fill then hits L2)
    double *red = smem;     // 32
    double *T = smem + 32;  // level table (levels 1..nlr in nrep bank-shifted replicas), T[pad] = 0, then R entries of w (exceptions)
    const int64_t R = (int64_t)1 << log2R;
    const int RL = 1 << (log2R + log2L);
    constexpr int K = BLOCK / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned padc = (unsigned)G.pad * 0x10001u;
    Chunk padq;
    padq.lo = make_uint4(padc, padc, padc, padc);
    padq.hi = padq.lo;
    bool first = true;
    for (;;) {
        __syncthreads();  // everybody is done with the previous tile's table (and has read cur_tile)
        if (threadIdx.x == 0) {
            // look-ahead only when every CTA has several tiles to go through: with few tiles per CTA (cells sharded over 8
            // GPUs: 319 tiles for 444 CTAs) reserving a second tile hoards work — half the CTAs would run two tiles in a row
            // while the others exit (measured: 0.15 -> 0.28 ms per product at 8 GPUs)
            if (ntiles >= 4 * (int64_t)gridDim.x) {
                cur_tile = first ? (long long)atomicAdd(&counters[0], 1u) : nxt_tile;
                nxt_tile = (long long)atomicAdd(&counters[0], 1u);
            } else {
                cur_tile = (long long)atomicAdd(&counters[0], 1u);
                nxt_tile = ntiles;
            }
        }
        first = false;
        __syncthreads();
        const int64_t t = cur_tile;
        if (t >= ntiles) break;
        if (nxt_tile < ntiles) {
            const char *nx = reinterpret_cast<const char *>(tlevA + (nxt_tile << (log2R + log2L)));
            for (int k = threadIdx.x * 128; k < RL * 8; k += BLOCK * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + k));
        }
        const int64_t row0 = t * G.rc;
        const double *tl = tlevA + (t << (log2R + log2L));  // this tile's level-major block
        // table fill: a thread owns a cell — w_i once, then its L table entries (all loads independent, coalesced per level)
        double part = 0.0;
        const int Lv = 1 << log2L;
        for (int il = threadIdx.x; il < G.rc; il += BLOCK) {
            const int64_t row = row0 + il;
            const double wv = (row < m) ? __ldg(w + row) : 0.0;
            T[G.wbase + il] = wv;
            part += wv;
            // ALL the loads of a batch of 16 levels first, then the products and the stores: ncu on the first version (unroll
            // 4) showed four serial L2 round trips per tile here — 12 % of the kernel's stall samples on its four DMULs
            for (int l0 = 0; l0 < Lv; l0 += 16) {
                double tv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) tv[j] = (l0 + j < Lv) ? __ldg(tl + ((l0 + j) << log2R) + il) : 0.0;  // zero beyond the last cell
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int l = l0 + j;
                    if (l < Lv) {
                        const double v = tv[j] * wv;
                        if (l < G.nlr) {
                            double *dst = T + l * G.levstride + il;
                            for (int r = 0; r < G.nrep; ++r) dst[r * G.strideA] = v;
                        } else {
                            T[G.baseB + ((l - G.nlr) << log2R) + il] = v;
                        }
                    }
                }
            }
        }
        if (threadIdx.x == 0) T[G.pad] = 0.0;
        // this warp's slice (loads issued before the barrier so that they overlap the table fill)
        const int g0 = __ldg(slices + t * (K + 1) + wid), g1 = __ldg(slices + t * (K + 1) + wid + 1);
        const int64_t cbeg = __ldg(gptr + t * n + g0), cend = __ldg(gptr + t * n + g1);
        // lane = chunk index mod 32 (bank-aware order, see the builder); 32-bit indices relative to the aligned start of the
        // slice and the FPD = 3 pipeline stages unrolled by hand — the first version rotated the stages through registers
        // (q[s] = q[s+1]) and did 64-bit index arithmetic: ncu counted 24 moves and ~12 carry instructions in the 112 of an
        // iteration (profiles/r04_spmv.md)
        const int64_t cbase = cbeg & ~(int64_t)31;
        const int lo = (int)(cbeg - cbase), hi = (int)(cend - cbase);
        const Chunk *pc = code + cbase + lane;
        const uint8_t *pm = meta + cbase + lane;
        int rel = lane;
        // NS stages in flight: 3 or 4 for the one-CTA-per-SM kernel (64-register budget), 2 for the three-CTAs-per-SM one (40)
        Chunk q0, q1, q2 = padq, q3 = padq;
        unsigned m0, m1, m2 = 0u, m3 = 0u;
        {
            const bool ok0 = rel >= lo && rel < hi, ok1 = rel + 32 < hi, ok2 = NS >= 3 && rel + 64 < hi, ok3 = NS >= 4 && rel + 96 < hi;
            q0 = ok0 ? ld_chunk(pc) : padq;
            m0 = ok0 ? (unsigned)__ldg(pm) : 0u;
            q1 = ok1 ? ld_chunk(pc + 32) : padq;
            m1 = ok1 ? (unsigned)__ldg(pm + 32) : 0u;
            if (NS >= 3) {
                q2 = ok2 ? ld_chunk(pc + 64) : padq;
                m2 = ok2 ? (unsigned)__ldg(pm + 64) : 0u;
            }
            if (NS >= 4) {
                q3 = ok3 ? ld_chunk(pc + 96) : padq;
                m3 = ok3 ? (unsigned)__ldg(pm + 96) : 0u;
            }
            pc += 32 * NS;
            pm += 32 * NS;
        }
        const double wsum = fblock_sum(part, red);  // publishes T
        if (threadIdx.x == 0) partial[t * (n + 1) + n] = wsum;
        double *prow = partial + t * (n + 1);
        int gbase = g0;
        double acc = 0.0;  // LAZY: this lane's share of the open segment; else the carried sum of the open segment
#define SVB_ADJ_STEP(QA, MA)                                                                                          \
    {                                                                                                                 \
        const unsigned mcur = (MA);                                                                                   \
        const unsigned bal = __ballot_sync(0xffffffffu, (mcur & 1u) != 0u);                                           \
        const int g = gbase + __popc(bal & lt);                                                                       \
        gbase += __popc(bal);                                                                                         \
        const double v = gather8(T, (QA).lo) + gather8(T, (QA).hi);                                                   \
        if (rel + 32 * NS < hi) { /* refill this stage with the chunk of iteration + NS */                            \
            (QA) = ld_chunk(pc);                                                                                      \
            (MA) = (unsigned)__ldg(pm);                                                                               \
        } else {                                                                                                      \
            (QA) = padq;                                                                                              \
            (MA) = 0u;                                                                                                \
        }                                                                                                             \
        pc += 32;                                                                                                     \
        pm += 32;                                                                                                     \
        if (LAZY) {                                                                                                   \
            const int nends = __popc(bal); /* segments that end in this iteration (warp-uniform) */                   \
            if (nends == 0) {                                                                                         \
                acc += v; /* the open segment continues: lane-private partial sums */                                 \
            } else {                                                                                                  \
                double r;                                                                                             \
                if (nends == 1) {                                                                                     \
                    const int b1 = __ffs(bal) - 1;                                                                    \
                    r = acc + ((lane <= b1) ? v : 0.0);                                                               \
                    _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);       \
                } else {                                                                                              \
                    double tot = acc;                                                                                 \
                    _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);   \
                    bool head0;                                                                                       \
                    r = seg_scan(v + (lane == 0 ? tot : 0.0), bal, lane, head0);                                      \
                }                                                                                                     \
                if (mcur & 1u) prow[g] = r * __ldg(inv + g);                                                          \
                acc = (lane > 31 - __clz(bal)) ? v : 0.0;                                                             \
            }                                                                                                         \
        } else {                                                                                                      \
            bool head0;                                                                                               \
            double vs = seg_scan(v, bal, lane, head0);                                                                \
            if (head0) vs += acc; /* acc = the carried sum of the segment that started in an earlier iteration */     \
            if (mcur & 1u) prow[g] = vs * __ldg(inv + g);                                                             \
            const double v31 = __shfl_sync(0xffffffffu, vs, 31);                                                      \
            acc = (bal >> 31) ? 0.0 : v31;                                                                            \
        }                                                                                                             \
        rel += 32;                                                                                                    \
    }
        while (rel - lane < hi) {
            SVB_ADJ_STEP(q0, m0)
            if (rel - lane >= hi) break;
            SVB_ADJ_STEP(q1, m1)
            if (NS >= 3) {
                if (rel - lane >= hi) break;
                SVB_ADJ_STEP(q2, m2)
            }
            if (NS >= 4) {
                if (rel - lane >= hi) break;
                SVB_ADJ_STEP(q3, m3)
            }
        }
#undef SVB_ADJ_STEP
    }
    // the last CTA to leave resets the counters for the next launch
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&counters[1], 1u);
        if (done == gridDim.x - 1) {
            counters[0] = 0u;
            counters[1] = 0u;
            __threadfence();
        }
    }
}

template <typename K>
static int fresident_grid(K kernel, size_t smem, int threads) {
    int per_sm = 0;
    SVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    return std::max(1, per_sm) * ctx().sm_count;
}

template <int BLOCK, bool BO, int MINB>
static void launch_fact_fwd(svb_operator_s *op, double alpha, const double *dx, double beta, double *dy, const double *coef,
                            double csign, const double *cvec) {
    svb_factored_s *f = op->fact;
    const size_t smem = (32 + (size_t)(f->f_nrep - 1) * f->f_stride + (size_t)op->n + 1) * sizeof(double);
    auto k = fwd_stream_kernel<BLOCK, BO, MINB>;
    SVB_CHECK(smem <= ctx().smem_optin, SVB_EDIM, "count-level operator: gene vector does not fit in shared memory");
    if (smem > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (f->fwd_grid == 0) {
        // one resident wave; fewer warps when the matrix is small (at least ~8 iterations of 32 chunks per warp)
        int grid = fresident_grid(k, smem, BLOCK);
        const int64_t want = std::max<int64_t>(1, f->f_chunks / (128 * (BLOCK / 32)));
        grid = (int)std::max<int64_t>(1, std::min<int64_t>(grid, want));
        const int NW = grid * (BLOCK / 32);
        SVB_CUDA(cudaMalloc((void **)&f->fwd_ranges, (size_t)(NW + 1) * sizeof(int64_t)));
        SVB_CUDA(cudaMalloc((void **)&f->fwd_rows, (size_t)(NW + 1) * sizeof(int64_t)));
        fact_warp_ranges_kernel<<<(NW + 256) / 256, 256, 0, ctx().stream>>>(f->f_rowptr, op->m, f->f_chunks, NW, f->fwd_ranges, f->fwd_rows);
        count_launch();
        SVB_LAUNCH_CHECK();
        f->fwd_grid = grid;
    }
    k<<<(unsigned)f->fwd_grid, BLOCK, smem, ctx().stream>>>((const Chunk *)f->f_code, f->f_meta, f->tlev, f->log2L, f->fwd_ranges,
                                                             f->fwd_rows, op->n, dx, f->inv, op->mu, alpha, beta, dy, coef, csign, cvec,
                                                             f->f_nrep, f->f_stride);
    SVB_LAUNCH_CHECK();
    if (f->excT) {
        fwd_exceptions_kernel<<<fgrid(op->m, 256, 148 * 8), 256, 0, ctx().stream>>>(f->excT->colptr, f->excT->rowidx, (const double *)f->excT->val,
                                                                                    op->m, dx, f->inv, alpha, dy);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
}

// adjoint CTA: 384 threads and a 64 KB table (R*L = 8192; three CTAs per SM), or -- SVB_FACT_LOG2R one larger -- 1024
// threads and a 128 KB table (R*L = 16384; one CTA per SM, segments twice as long)
// SVB_ADJ_TWO=1 (experiment): tiles of R*L = 8192 entries WITH replica tables (94.5 KB), two 512-thread CTAs per SM — one CTA
// fills its table while the other streams
static inline bool adj_two_mode() {
    static const bool two = getenv("SVB_ADJ_TWO") && atoi(getenv("SVB_ADJ_TWO")) != 0;
    return two;
}
static inline int adj_block_of(const svb_factored_s *f) {
    if (f->log2R + f->log2L >= 14) return 1024;
    return (adj_two_mode() && f->a_nrep > 1) ? 512 : 384;
}

void fact_fwd(svb_operator_s *op, double alpha, const double *dx, double beta, double *dy, const double *coef, double csign,
              const double *cvec) {
    const bool bo = op->fact->f_cshift == 3;  // codes are byte offsets
    // 64 registers per thread (4 x 256 or 2 x 512 threads per SM): 48 registers spill, and the kernel is bound by the
    // shared-memory pipe / issue slots, not by occupancy (measured 5 vs 4 CTAs per SM: within noise)
    if (((size_t)(op->fact->f_nrep - 1) * op->fact->f_stride + (size_t)op->n) * 8 > 24 * 1024) {
        // two CTAs per SM: 512 threads (64 registers, 32 warps per SM) or 448 (72 registers, 28 warps); SVB_FWD_BLOCK selects
        static const int blk = getenv("SVB_FWD_BLOCK") ? atoi(getenv("SVB_FWD_BLOCK")) : 448;
        if (blk == 448) {
            if (bo) launch_fact_fwd<448, true, 2>(op, alpha, dx, beta, dy, coef, csign, cvec);
            else launch_fact_fwd<448, false, 2>(op, alpha, dx, beta, dy, coef, csign, cvec);
        } else {
            if (bo) launch_fact_fwd<512, true, 2>(op, alpha, dx, beta, dy, coef, csign, cvec);
            else launch_fact_fwd<512, false, 2>(op, alpha, dx, beta, dy, coef, csign, cvec);
        }
    } else {
        if (bo) launch_fact_fwd<256, true, 3>(op, alpha, dx, beta, dy, coef, csign, cvec);
        else launch_fact_fwd<256, false, 3>(op, alpha, dx, beta, dy, coef, csign, cvec);
    }
}

void fact_adj_stage1(svb_operator_s *op, const double *dx) {
    svb_factored_s *f = op->fact;
    const size_t smem = (32 + (size_t)f->a_tabsize) * sizeof(double);
    const int block = adj_block_of(f);
    const AdjGeom G{f->a_nlr, f->a_nrep, f->a_strideA, f->a_levstride, f->a_baseB, f->a_pad, f->a_wbase, (int)f->Rc};
    static const int stages = getenv("SVB_ADJ_STAGES") ? atoi(getenv("SVB_ADJ_STAGES")) : 2;
    static const bool lazy2 = getenv("SVB_ADJ_LAZY") && atoi(getenv("SVB_ADJ_LAZY")) != 0;
    auto k = (block == 1024) ? (stages >= 3 ? adj_stream_kernel<1024, 1, true, 3> : adj_stream_kernel<1024, 1, true, 2>)
             : (block == 512) ? (lazy2 ? adj_stream_kernel<512, 2, true, 2> : adj_stream_kernel<512, 2, false, 2>)
                              : adj_stream_kernel<384, 3, false, 2>;  // 3 x 384 threads: 56 registers (512 threads: 40, spills)
    if (smem > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (f->adj_grid == 0) f->adj_grid = (int)std::max<int64_t>(1, std::min<int64_t>(fresident_grid(k, smem, block), f->ntiles));
    k<<<(unsigned)f->adj_grid, block, smem, ctx().stream>>>(f->a_gptr, (const Chunk *)f->a_code, f->a_meta, f->tlevA, f->log2L, f->log2R,
                                                                 op->m, op->n, f->ntiles, dx, f->inv, f->partial, f->a_slices, f->counters, G);
    SVB_LAUNCH_CHECK();
    if (f->exc && f->e_nseg > 0) {  // the exception side sums, one CTA per segment of <= 4096 entries (added by the reduce kernel)
        adj_exceptions_kernel<<<(unsigned)f->e_nseg, 256, 0, ctx().stream>>>(f->exc->colptr, f->exc->rowidx, (const double *)f->exc->val, dx,
                                                                             f->e_segptr, op->n, f->e_segsum);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
}

// algorithmic bytes of one product with THIS layout's widths: 2 B per stored entry and 1 B per chunk of 8 (pads are not
// counted), the per-cell tables, the pointers and the vectors
double fact_fwd_bytes(const svb_operator_s *op) {
    const svb_factored_s *f = op->fact;
    // exception entries: 12 B each (Float64 value + Int32 gene) + the cell pointers of the side matrix
    return (2.0 + 1.0 / FCH) * (double)f->nnz_main + 12.0 * (double)f->nnz_exc + (f->nnz_exc ? 8.0 * op->m : 0.0) +
           (double)op->m * (8.0 * f->L + 8.0) + 24.0 * (double)op->n;
}
double fact_adj_bytes(const svb_operator_s *op) {
    const svb_factored_s *f = op->fact;
    const double nseg = (double)f->ntiles * (double)op->n;
    return (2.0 + 1.0 / FCH) * (double)f->nnz_main + 12.0 * (double)f->nnz_exc + (double)op->m * (8.0 * f->L + 8.0) + nseg * (8.0 + 8.0) +
           24.0 * (double)op->n;
}

// ---------------------------------------------------------------------------------------------
// construction
// ---------------------------------------------------------------------------------------------
namespace {
// a non-owning view of a CSC matrix whose value array is replaced (the level bytes)
struct MatrixView {
    svb_matrix_s v;
    MatrixView(const svb_matrix_s *a, void *val) {
        v.nrow = a->nrow;
        v.ncol = a->ncol;
        v.nnz = a->nnz;
        v.vtype = a->vtype;
        v.colptr = a->colptr;
        v.rowidx = a->rowidx;
        v.val = val;
    }
    ~MatrixView() { v.colptr = nullptr; v.rowidx = nullptr; v.val = nullptr; }
};
}  // namespace

static void build_factored(svb_operator_s *op, const svb_matrix_s *a, const int64_t *h_libsize, double sf, const double *h_mean,
                           const double *h_var, double scale_max, int levels, double *mu_out) {
    Context &C = ctx();
    cudaStream_t st = C.stream;
    const int64_t m = a->nrow, n = a->ncol, nnz = a->nnz;
    auto *f = new svb_factored_s();
    op->fact = f;
    // SVB_FACT_TIMING=1: per-phase build times (stream-synchronised) on stderr
    const bool timing = getenv("SVB_FACT_TIMING") != nullptr;
    auto tlast = std::chrono::steady_clock::now();
    auto tick = [&](const char *what) {
        if (!timing) return;
        cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[svb counts build] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tlast).count());
        tlast = now;
    };

    // ---- number of levels ----------------------------------------------------------------------------
    DevBuf<long long> d_lib((size_t)std::max<int64_t>(m, 1));
    SVB_CUDA(cudaMemcpyAsync(d_lib.p, h_libsize, (size_t)m * 8, cudaMemcpyHostToDevice, st));
    int L = levels;
    if (L == 0) {
        DevBuf<unsigned long long> d_h(4);
        SVB_CUDA(cudaMemsetAsync(d_h.p, 0, 4 * sizeof(unsigned long long), st));
        if (nnz > 0) {
            fact_hist_kernel<<<fgrid(nnz), 256, 0, st>>>((const int32_t *)a->val, nnz, d_h.p);
            count_launch();
            SVB_LAUNCH_CHECK();
        }
        double h[5] = {0, 0, 0, 0, 0};
        {
            unsigned long long hh[4];
            SVB_CUDA(cudaMemcpyAsync(hh, d_h.p, sizeof(hh), cudaMemcpyDeviceToHost, st));
            SVB_CUDA(cudaStreamSynchronize(st));
            for (int i = 0; i < 4; ++i) h[i] = (double)hh[i];
            h[4] = (double)nnz;
        }
        if (C.nranks > 1) {  // every rank must choose the same L (tables are per rank, but keep the ranks in lock-step)
            DevBuf<double> d5(5);
            SVB_CUDA(cudaMemcpyAsync(d5.p, h, sizeof(h), cudaMemcpyHostToDevice, st));
            comm_allreduce_dev(d5.p, 5);
            SVB_CUDA(cudaMemcpyAsync(h, d5.p, sizeof(h), cudaMemcpyDeviceToHost, st));
            SVB_CUDA(cudaStreamSynchronize(st));
        }
        // smallest L that leaves at most 2 % of the entries to the exception chunks (17 B and a whole lane slot per entry
        // there, 2.1 B here; doubling L halves the cells per adjoint tile, which costs more than 2 % of exceptions)
        const int cand[4] = {4, 8, 16, 32};
        L = 32;
        for (int i = 0; i < 4; ++i)
            if (h[i] <= 0.02 * h[4]) { L = cand[i]; break; }
    }
    int log2L = 0;
    while ((1 << log2L) < L) ++log2L;
    f->L = L;
    f->log2L = log2L;
    // Adjoint tiles. The shared-memory table has room for R*L entries: 16384 (128 KB + replicas, one 1024-thread CTA per SM;
    // measured 0.67 vs 0.76 ms per product at C3) from 256 cells per SM on, else 8192 (64 KB, three 384-thread CTAs per SM);
    // small inputs get one small tile. A tile holds Rc <= R cells, chosen so that the tiles fill WHOLE rounds of the grid:
    // k = ceil(m / (R * SMs)) rounds of Rc = m / (k * SMs) cells (rounded up to 16: the bank of an entry stays cell mod 16).
    // With Rc = R a shard of 163,266 cells (C3 on 8 GPUs) was 160 tiles on 148 SMs: two rounds for the work of 1.08.
    const bool big = m >= (int64_t)256 * C.sm_count;
    int log2R = big ? 14 - log2L : 13 - log2L;
    const char *envr = getenv("SVB_FACT_LOG2R");
    if (envr) log2R = std::max(5, std::min(atoi(envr), 14 - log2L));
    while (log2R > 5 && (1ll << (log2R - 1)) >= m) --log2R;
    f->log2R = log2R;
    f->R = 1ll << log2R;
    f->Rc = f->R;
    static const bool whole_rounds = !(getenv("SVB_FACT_ROUNDS") && atoi(getenv("SVB_FACT_ROUNDS")) == 0);
    if (whole_rounds && big && log2R + log2L >= 14 && log2R >= 6) {
        const int64_t slots = C.sm_count;
        const int64_t k = (m + f->R * slots - 1) / (f->R * slots);
        const int64_t rc = ((m + k * slots - 1) / (k * slots) + 15) & ~(int64_t)15;
        f->Rc = std::max<int64_t>(16, std::min<int64_t>(rc, f->R));
    }
    f->ntiles = std::max<int64_t>(1, (m + f->Rc - 1) / f->Rc);
    // ---- bank-shifted replicas of the gathered tables (see fact_assign_kernel) ------------------------------------------
    static const bool replicas = !(getenv("SVB_FACT_REPLICAS") && atoi(getenv("SVB_FACT_REPLICAS")) == 0);
    {
        // forward: xs replicas while the byte-offset codes still fit in 16 bits; replica r starts at entry r*stride, stride = 5 mod 16
        int stride = (int)n + 1;
        while ((stride & 15) != 5) ++stride;
        f->f_stride = stride;
        f->f_nrep = 1;
        if (replicas && n <= 8190) f->f_nrep = (int)std::max<int64_t>(1, std::min<int64_t>(4, (8191 - n) / stride + 1));
        // adjoint: only the one-CTA-per-SM kernel (R*L = 16384) has the shared memory for it: levels 1..2 in 4 replicas
        const int RL = 1 << (log2R + log2L);
        // the largest replica configuration that fits the shared memory of an SM (simulated passes per set at C3, L = 16,
        // R = 1024: {levels 1-2 x 4: 1.10, 1-2 x 3: 1.19, level 1 x 4: 1.19, 1-2 x 2: ~1.3}; no replicas: 1.55)
        const int cand[6][2] = {{2, 4}, {2, 3}, {1, 4}, {2, 2}, {1, 3}, {1, 2}};
        auto set_geom = [&](int nlr, int nrep) {
            f->a_nlr = nlr;
            f->a_nrep = nrep;
            f->a_strideA = nrep > 1 ? (int)f->R + 5 : 0;                              // replica r of a level: 5r banks later
            f->a_levstride = nrep > 1 ? (nrep * f->a_strideA + 15) & ~15 : 0;          // levels 16-aligned: replica 0 keeps bank = cell mod 16
            f->a_baseB = nlr * f->a_levstride;
            f->a_pad = f->a_baseB + RL - (nlr << log2R);
            f->a_wbase = f->a_pad + 1;
            f->a_tabsize = f->a_wbase + (int)f->R;
            const size_t budget = (log2R + log2L >= 14) ? (size_t)C.smem_optin : (size_t)112 * 1024;  // two CTAs per SM below 16384 entries
            return (size_t)(32 + f->a_tabsize) * sizeof(double) <= budget && f->a_tabsize <= 65535;
        };
        bool done = false;
        if (replicas && (log2R + log2L >= 14 || (adj_two_mode() && log2R + log2L == 13)) && log2R >= 4)
            for (int i = 0; i < 6 && !done; ++i)
                if (cand[i][0] <= L) done = set_geom(cand[i][0], cand[i][1]);
        if (!done) set_geom(0, 1);
    }

    if (timing)
        fprintf(stderr, "[svb counts build] L %d log2R %d tiles %lld | fwd replicas %d stride %d | adj nlr %d nrep %d strideA %d levstride %d baseB %d pad %d table %d entries (%zu B of %zu)\n",
                L, log2R, (long long)f->ntiles, f->f_nrep, f->f_stride, f->a_nlr, f->a_nrep, f->a_strideA, f->a_levstride, f->a_baseB,
                f->a_pad, f->a_tabsize, (size_t)(32 + f->a_tabsize) * sizeof(double), (size_t)C.smem_optin);
    tick("libsize upload + levels");
    // ---- per-cell level tables ---------------------------------------------------------------------
    SVB_CUDA(cudaMalloc((void **)&f->tlev, (size_t)std::max<int64_t>(m << log2L, 1) * sizeof(double)));
    SVB_CUDA(cudaMalloc((void **)&f->tlevA, (size_t)(f->ntiles << (log2R + log2L)) * sizeof(double)));
    SVB_CUDA(cudaMemsetAsync(f->tlevA, 0, (size_t)(f->ntiles << (log2R + log2L)) * sizeof(double), st));
    fact_tlev_kernel<<<fgrid(m << log2L), 256, 0, st>>>(d_lib.p, m, log2L, log2R, f->Rc, sf, f->tlev, f->tlevA);
    count_launch();
    SVB_LAUNCH_CHECK();

    tick("level tables");
    // ---- per-gene moments (given, or two parallel passes over the counts), then scale, centre, clip ------
    DevBuf<double> d_mean((size_t)n), d_var((size_t)n), d_sd((size_t)n), d_cap((size_t)n);
    DevBuf<int> d_bad(1);
    SVB_CUDA(cudaMalloc((void **)&op->mu, (size_t)n * sizeof(double)));
    SVB_CUDA(cudaMalloc((void **)&f->inv, (size_t)n * sizeof(double)));
    if (h_mean) {
        SVB_CUDA(cudaMemcpyAsync(d_mean.p, h_mean, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        SVB_CUDA(cudaMemcpyAsync(d_var.p, h_var, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    } else {
        DevBuf<double> d_sum((size_t)n + 1), d_ss((size_t)n);
        const double mloc = (double)m;
        SVB_CUDA(cudaMemcpyAsync(d_sum.p + n, &mloc, 8, cudaMemcpyHostToDevice, st));
        fact_moments_sum_kernel<<<(unsigned)n, 512, 0, st>>>(a->colptr, a->rowidx, (const int32_t *)a->val, log2L, f->tlev, d_lib.p, sf,
                                                             d_sum.p);
        count_launch();
        SVB_LAUNCH_CHECK();
        if (C.nranks > 1) comm_allreduce_dev(d_sum.p, n + 1);  // cells are sharded: sums and the cell count over all ranks
        fact_moments_ss_kernel<<<(unsigned)n, 512, 0, st>>>(a->colptr, a->rowidx, (const int32_t *)a->val, log2L, f->tlev, d_lib.p, sf, m,
                                                            d_sum.p, d_sum.p + n, d_mean.p, d_ss.p);
        count_launch();
        SVB_LAUNCH_CHECK();
        if (C.nranks > 1) comm_allreduce_dev(d_ss.p, n);
        fact_moments_var_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_ss.p, d_sum.p + n, n, d_var.p);
        count_launch();
        SVB_LAUNCH_CHECK();
        SVB_CUDA(cudaStreamSynchronize(st));  // d_sum / d_ss go out of scope
    }
    SVB_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
    fact_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_mean.p, d_var.p, n, scale_max, d_sd.p, op->mu, d_cap.p, f->inv, d_bad.p);
    count_launch();
    SVB_LAUNCH_CHECK();
    int bad = 0;
    SVB_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (mu_out) SVB_CUDA(cudaMemcpyAsync(mu_out, op->mu, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_CHECK(!bad, SVB_EARG, "svb_operator_create_counts: a gene has zero (or non-finite) variance");

    tick("moments + prepare");
    // ---- classify every entry (CSC order) -------------------------------------------------------------
    DevBuf<uint8_t> lvl((size_t)std::max<int64_t>(nnz, 1));
    DevBuf<int64_t> exoff((size_t)(n * FSL + 1));  // exceptions per (gene, slice) -> offsets
    SVB_CUDA(cudaMemsetAsync(exoff.p, 0, (size_t)(n * FSL + 1) * sizeof(int64_t), st));
    {
        dim3 grid((unsigned)n, FSL / 8);
        fact_classify_kernel<<<grid, 256, 0, st>>>(a->colptr, a->rowidx, (const int32_t *)a->val, log2L, f->tlev, d_sd.p, d_cap.p, lvl.p,
                                                   exoff.p);
        count_launch();
        SVB_LAUNCH_CHECK();
    }
    exclusive_scan_i64(exoff.p, n * FSL + 1, st);
    int64_t nexc = 0;
    SVB_CUDA(cudaMemcpyAsync(&nexc, exoff.p + n * FSL, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    f->nnz_exc = nexc;
    f->nnz_main = nnz - nexc;

    tick("classify + scan");
    // ---- the exceptions with their exact value (times sd), gene-major (CSC) --------------------------------
    struct Owned {
        svb_matrix_s *p = nullptr;
        ~Owned() { delete p; }
    } e;
    if (nexc > 0) {
        e.p = matrix_alloc(m, n, nexc, SVB_F64);
        launch_strided_copy(exoff.p, FSL, n, e.p->colptr, st);
        SVB_CUDA(cudaMemcpyAsync(e.p->colptr + n, exoff.p + n * FSL, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        dim3 grid((unsigned)n, FSL / 8);
        fact_exception_fill_kernel<<<grid, 256, 0, st>>>(a->colptr, a->rowidx, (const int32_t *)a->val, lvl.p, d_lib.p, sf, d_sd.p, d_cap.p,
                                                     exoff.p, e.p->rowidx, (double *)e.p->val);
        count_launch();
        SVB_LAUNCH_CHECK();
    }

    tick("exceptions (CSC)");
    // ---- adjoint layout ----------------------------------------------------------------------------------
    {
        const int64_t nseg = f->ntiles * n;
        DevBuf<int64_t> startpos((size_t)((f->ntiles + 1) * n + 1)), estart;  // estart stays empty: no exception chunks
        tile_bounds(a, f->Rc, f->ntiles, startpos.p);

        SVB_CUDA(cudaMalloc((void **)&f->a_gptr, (size_t)(nseg + 1) * sizeof(int64_t)));
        fact_seg_count_kernel<<<fgrid(nseg * 8), 256, 0, st>>>(startpos.p, lvl.p, estart.p, nseg, n, f->a_gptr);
        count_launch();
        SVB_LAUNCH_CHECK();
        exclusive_scan_i64(f->a_gptr, nseg + 1, st);
        SVB_CUDA(cudaMemcpyAsync(&f->a_chunks, f->a_gptr + nseg, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        SVB_CUDA(cudaMalloc((void **)&f->a_code, (size_t)std::max<int64_t>(f->a_chunks, 1) * FCH * 2));
        SVB_CUDA(cudaMalloc((void **)&f->a_meta, (size_t)std::max<int64_t>(f->a_chunks, 1)));
        fact_seg_fill_kernel<<<fgrid(nseg * 8), 256, 0, st>>>(startpos.p, a->rowidx, lvl.p, nseg, n, log2R, f->Rc, log2L, f->a_gptr, estart.p,
                                                              nullptr, nullptr, (uint16_t *)f->a_code, f->a_meta);
        count_launch();
        SVB_LAUNCH_CHECK();
        tick("  adjoint: segments + placement");
        static const bool stats = getenv("SVB_FACT_STATS") != nullptr;
        if (f->a_nrep > 1 || stats) {
            const AssignGeom G{1, f->a_nrep, f->a_strideA ? f->a_strideA : 5, 1 << (log2R + log2L), log2R, f->a_nlr, f->a_baseB, f->a_pad, f->a_levstride};
            f->a_passes = run_assign((uint16_t *)f->a_code, f->a_meta, f->a_chunks, G, st);
            tick("  adjoint: replica assignment");
        }
        const int K = adj_block_of(f) / 32;
        SVB_CUDA(cudaMalloc((void **)&f->a_slices, (size_t)f->ntiles * (K + 1) * sizeof(int32_t)));
        static const int slice_alpha = getenv("SVB_ADJ_SLICE_ALPHA") ? atoi(getenv("SVB_ADJ_SLICE_ALPHA")) : 8;
        fact_slices_kernel<<<(unsigned)((f->ntiles * (K + 1) + 255) / 256), 256, 0, st>>>(f->a_gptr, f->ntiles, n, K, slice_alpha, f->a_slices);
        SVB_CUDA(cudaMalloc((void **)&f->counters, 2 * sizeof(unsigned int)));
        SVB_CUDA(cudaMemsetAsync(f->counters, 0, 2 * sizeof(unsigned int), st));
        count_launch();
        SVB_LAUNCH_CHECK();
        SVB_CUDA(cudaStreamSynchronize(st));
    }

    tick("adjoint layout");
    // ---- forward layout: CSR by cell of (gene, level), then level grouping -------------------------------------
    {
        MatrixView view(a, lvl.p);
        TileCSC<uint8_t> tc;
        DevBuf<int64_t> rowptr;
        DevBuf<uint16_t> ridx;
        DevBuf<uint8_t> rlvl;
        int tl2 = 12;
        while (tl2 > 8 && (1ll << (tl2 - 1)) >= m) --tl2;
        build_tilecsc<uint8_t, uint8_t>(&view.v, tl2, tc);
        csr_from_tilecsc<uint8_t, uint16_t>(tc, &view.v, rowptr, ridx, rlvl);
        tc.gptr.release();
        tc.rloc.release();
        tc.aval.release();
        tick("forward: CSR of (gene, level)");
        SVB_CUDA(cudaMalloc((void **)&f->f_rowptr, (size_t)(m + 1) * sizeof(int64_t)));
        DevBuf<uint16_t> gend((size_t)std::max<int64_t>(m * L, 1));
        const unsigned gw = (unsigned)std::max<int64_t>(1, std::min<int64_t>((m + 7) / 8, 148 * 16));
        fact_row_hist_kernel<<<gw, 256, 0, st>>>(rowptr.p, rlvl.p, m, L, gend.p, f->f_rowptr);
        count_launch();
        SVB_LAUNCH_CHECK();
        exclusive_scan_i64(f->f_rowptr, m + 1, st);
        SVB_CUDA(cudaMemcpyAsync(&f->f_chunks, f->f_rowptr + m, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        SVB_CUDA(cudaMalloc((void **)&f->f_code, (size_t)std::max<int64_t>(f->f_chunks, 1) * FCH * 2));
        SVB_CUDA(cudaMalloc((void **)&f->f_meta, (size_t)std::max<int64_t>(f->f_chunks, 1)));
        f->f_cshift = (n <= 8190) ? 3 : 0;  // the codes of the forward stream are byte offsets when they fit in 16 bits
        fact_row_place_kernel<<<gw, 256, (size_t)8 * 2 * L * 16 * sizeof(int), st>>>(rowptr.p, ridx.p, rlvl.p, m, L, (int)n, f->f_cshift, f->f_rowptr, gend.p,
                                                  nullptr, nullptr, nullptr, (uint16_t *)f->f_code, f->f_meta);
        count_launch();
        SVB_LAUNCH_CHECK();
        tick("  forward: placement");
        static const bool stats = getenv("SVB_FACT_STATS") != nullptr;
        if (f->f_cshift == 3 && (f->f_nrep > 1 || stats)) {
            const AssignGeom G{0, f->f_nrep, f->f_stride, (int)n, 0, 0, 0, 0, 0};
            f->f_passes = run_assign((uint16_t *)f->f_code, f->f_meta, f->f_chunks, G, st);
            tick("  forward: replica assignment");
        }
        SVB_CUDA(cudaStreamSynchronize(st));
    }

    tick("forward: grouping + placement");
    if (e.p) {  // the forward's view: the same entries cell-major (stable device transpose), then both are kept
        f->excT = matrix_transpose(e.p);
        f->exc = e.p;
        e.p = nullptr;
        SVB_CUDA(cudaMalloc((void **)&f->e_segptr, (size_t)(n + 1) * sizeof(int64_t)));
        exc_segcount_kernel<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(f->exc->colptr, n, f->e_segptr);
        count_launch();
        SVB_LAUNCH_CHECK();
        exclusive_scan_i64(f->e_segptr, n + 1, st);
        SVB_CUDA(cudaMemcpyAsync(&f->e_nseg, f->e_segptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        SVB_CUDA(cudaStreamSynchronize(st));
        SVB_CUDA(cudaMalloc((void **)&f->e_segsum, (size_t)std::max<int64_t>(f->e_nseg, 1) * sizeof(double)));
        tick("exception side matrices");
    }
    SVB_CUDA(cudaMalloc((void **)&f->partial, (size_t)f->ntiles * (n + 1) * sizeof(double)));
    SVB_CUDA(cudaMalloc((void **)&op->tmp, (size_t)std::max(m, n) * sizeof(double)));
    SVB_CUDA(cudaMalloc((void **)&op->scal, 8 * sizeof(double)));
    SVB_CUDA(cudaStreamSynchronize(st));
    tick("partial / scratch allocation");
}

}  // namespace svb

extern "C" {

int svb_operator_create_counts(svb_matrix_t counts, const int64_t *libsize, double scale_factor, const double *mean,
                               const double *var, double scale_max, int levels, double *mu_out, svb_operator_t *out) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(counts && libsize && out, SVB_EARG, "svb_operator_create_counts: null argument");
    SVB_CHECK((mean == nullptr) == (var == nullptr), SVB_EARG, "svb_operator_create_counts: give both moments or neither");
    SVB_CHECK(counts->vtype == SVB_I32, SVB_EARG, "svb_operator_create_counts: the matrix must hold integer counts");
    SVB_CHECK(counts->nrow >= 1 && counts->ncol >= 1, SVB_EDIM, "svb_operator_create_counts: empty operator");
    SVB_CHECK(counts->ncol <= 65534, SVB_EDIM, "svb_operator_create_counts: at most 65534 genes (16-bit codes); use svb_operator_create");
    SVB_CHECK(levels == 0 || levels == 4 || levels == 8 || levels == 16 || levels == 32, SVB_EARG,
              "svb_operator_create_counts: levels must be 0 (auto), 4, 8, 16 or 32");
    SVB_CHECK(scale_factor > 0.0, SVB_EARG, "svb_operator_create_counts: scale_factor must be positive");
    auto *op = new svb_operator_s();
    const auto t_begin = std::chrono::steady_clock::now();
    try {
        op->m = counts->nrow;
        op->n = counts->ncol;
        op->nnz = counts->nnz;
        op->vbytes = 0;
        op->ibytes = 2;
        build_factored(op, counts, libsize, scale_factor, mean, var, scale_max, levels, mu_out);
    } catch (...) {
        delete op;
        throw;
    }
    if (getenv("SVB_FACT_TIMING"))
        fprintf(stderr, "[svb counts build] %-28s %8.3f ms\n", "TOTAL (incl. temporaries freed)",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    *out = op;
    SVB_API_END
}

int svb_operator_counts_info(svb_operator_t op, int *levels, int64_t *tile_cells, int64_t *nnz_coded, int64_t *nnz_exception,
                             int64_t *fwd_chunks, int64_t *adj_chunks) {
    SVB_API_BEGIN
    SVB_CHECK(op && op->fact, SVB_EARG, "svb_operator_counts_info: not a count-level operator");
    const svb_factored_s *f = op->fact;
    if (levels) *levels = f->L;
    if (tile_cells) *tile_cells = f->Rc;
    if (nnz_coded) *nnz_coded = f->nnz_main;
    if (nnz_exception) *nnz_exception = f->nnz_exc;
    if (fwd_chunks) *fwd_chunks = f->f_chunks;
    if (adj_chunks) *adj_chunks = f->a_chunks;
    SVB_API_END
}

int svb_operator_counts_stream(svb_operator_t op, int adjoint, uint16_t *code, uint8_t *meta) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(op && op->fact, SVB_EARG, "svb_operator_counts_stream: not a count-level operator");
    const svb_factored_s *f = op->fact;
    const int64_t nch = adjoint ? f->a_chunks : f->f_chunks;
    cudaStream_t st = ctx().stream;
    if (code) SVB_CUDA(cudaMemcpyAsync(code, adjoint ? f->a_code : f->f_code, (size_t)nch * FCH * 2, cudaMemcpyDeviceToHost, st));
    if (meta) SVB_CUDA(cudaMemcpyAsync(meta, adjoint ? f->a_meta : f->f_meta, (size_t)nch, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    SVB_API_END
}

int svb_operator_counts_layout(svb_operator_t op, int *fwd_replicas, double *fwd_passes, int *adj_replicas,
                               int *adj_replicated_levels, double *adj_passes) {
    SVB_API_BEGIN
    SVB_CHECK(op && op->fact, SVB_EARG, "svb_operator_counts_layout: not a count-level operator");
    const svb_factored_s *f = op->fact;
    if (fwd_replicas) *fwd_replicas = f->f_nrep;
    if (fwd_passes) *fwd_passes = f->f_passes;
    if (adj_replicas) *adj_replicas = f->a_nrep;
    if (adj_replicated_levels) *adj_replicated_levels = f->a_nlr;
    if (adj_passes) *adj_passes = f->a_passes;
    SVB_API_END
}

}  // extern "C"

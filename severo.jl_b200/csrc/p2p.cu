// p2p.cu — one-shot allreduce over NVLink peer memory for the small per-step messages of the sharded
// Lanczos recurrence (S'w partial: n doubles; reorthogonalisation coefficients: <= w doubles; a norm: 1).
// These collectives are latency-bound (16 KB), not bandwidth-bound: NCCL costs ~25-30 us per call, three
// times per Lanczos step. Here every rank owns a "mailbox" in its HBM that all peers map through CUDA IPC;
// one single-CTA kernel stores the rank's vector straight into every peer's mailbox (remote st.global over
// NVLink), publishes an epoch flag (fence.sys + flag store), waits for the peers' flags and sums the slots
// in rank order — the sum is bitwise identical on every rank, which keeps the ranks' host control flow in
// lock-step. Slots and flags are double-buffered by epoch parity (a rank cannot be two epochs ahead because
// finishing epoch e needs every peer's epoch-e data, which a peer writes only after it finished e-1).
// Mailbox reads bypass L1 (ld.cg): L2 is the coherence point for peer writes. Every spin is bounded.
#define SVB_NO_ALLOC_MACROS
#include "svb_internal.h"
#include "p2p.cuh"

#include <cstring>

namespace svb {

struct P2PState {
    bool ready = false;
    int nranks = 1, rank = 0;
    Mailbox *mine = nullptr;
    Mailbox *peers_host[P2P_MAX_RANKS] = {nullptr};
    Mailbox **peers_dev = nullptr;  // device array of the mapped peer mailboxes (own entry = local pointer)
    unsigned long long epoch = 0;
    unsigned int *counter = nullptr;  // last-block counter for fused producers
    bool ipc = true;                  // peers mapped through CUDA IPC (one process per GPU); false: plain peer access (one process, N GPUs)
};
static P2PState &p2p_of_ctx() {
    Context &C = ctx();
    if (!C.p2p_state) C.p2p_state = new P2PState();
    return *static_cast<P2PState *>(C.p2p_state);
}
#define g_p2p (p2p_of_ctx())

// buf[0..n) <- sum over ranks of buf (in place). One CTA.
__global__ void __launch_bounds__(1024) p2p_allreduce_kernel(Mailbox *const *__restrict__ peers, Mailbox *__restrict__ mine,
                                                            int nranks, int rank, unsigned long long epoch,
                                                            double *__restrict__ buf, int n, long long timeout_cycles) {
    const int par = (int)(epoch & 1ull);
    const bool dead = *((volatile int *)&mine->error) != 0;  // a previous exchange timed out: do not spin again
    // 1. my contribution into slot [par][rank] of every mailbox (remote stores for the peers)
    for (int q = 0; q < nranks; ++q) {
        double *dst = peers[q]->slots[par][rank];
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: flag[par][rank] = epoch in every mailbox
    if (threadIdx.x < nranks) st_flag_sys(&peers[threadIdx.x]->flags[par][rank], epoch);
    // 3. wait for every peer's flag in my mailbox (bounded)
    if (threadIdx.x < nranks) {
        const long long t0 = clock64();
        while (ld_flag_sys(&mine->flags[par][threadIdx.x]) < epoch) {
            if (dead || clock64() - t0 > timeout_cycles) {
                mine->error = 1;
                break;
            }
        }
    }
    __syncthreads();
    // 4. sum the slots in rank order (same order on every rank -> same bits everywhere)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < nranks; ++q) s += __ldcg(&mine->slots[par][q][i]);
        buf[i] = s;
    }
}

bool p2p_ready() { return g_p2p.ready; }

bool p2p_next_ctx(int64_t n, P2PCtx *out) {
    if (!g_p2p.ready || n > P2P_CAP) return false;
    ++g_p2p.epoch;
    out->peers = g_p2p.peers_dev;
    out->mine = g_p2p.mine;
    out->nranks = g_p2p.nranks;
    out->rank = g_p2p.rank;
    out->epoch = g_p2p.epoch;
    out->timeout_cycles = 20000000000ll;  // ~10 s
    out->counter = g_p2p.counter;
    return true;
}

bool p2p_allreduce(double *dbuf, int64_t n) {
    if (!g_p2p.ready || n > P2P_CAP) return false;
    Context &C = ctx();
    ++g_p2p.epoch;
    KTimer kt(SVB_K_COMM, 8.0 * n, 1);
    const int threads = n >= 1024 ? 1024 : (n >= 256 ? 256 : 64);
    p2p_allreduce_kernel<<<1, threads, 0, C.stream>>>(g_p2p.peers_dev, g_p2p.mine, g_p2p.nranks, g_p2p.rank, g_p2p.epoch, dbuf,
                                                      (int)n, 20000000000ll /* ~10 s */);
    SVB_LAUNCH_CHECK();
    return true;
}

int p2p_error() {
    if (!g_p2p.ready) return 0;
    int e = 0;
    cudaMemcpyAsync(&e, &g_p2p.mine->error, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream);
    cudaStreamSynchronize(ctx().stream);
    return e;
}

// Called by svb_comm_init after the NCCL communicator exists (it carries the IPC handles).
void p2p_setup(int nranks, int rank) {
    g_p2p = P2PState();
    const char *env = getenv("SVB_P2P");
    if (env && atoi(env) == 0) return;
    if (nranks < 2 || nranks > P2P_MAX_RANKS) return;
    Context &C = ctx();
    // recorded first: p2p_teardown() walks peers_host[0..nranks) and must see every handle opened below, also when the
    // setup bails out half way (a failed cudaIpcOpenMemHandle, ranks that do not agree) — round-1 advice
    g_p2p.nranks = nranks;
    g_p2p.rank = rank;
    try {
        SVB_CUDA(cudaMalloc((void **)&g_p2p.mine, sizeof(Mailbox)));  // plain cudaMalloc: pool memory has no legacy IPC handle
        SVB_CUDA(cudaMemsetAsync(g_p2p.mine, 0, sizeof(Mailbox), C.stream));
        SVB_CUDA(cudaStreamSynchronize(C.stream));
        cudaIpcMemHandle_t h;
        SVB_CUDA(cudaIpcGetMemHandle(&h, g_p2p.mine));
        // all-gather the 64-byte handles through the NCCL allreduce: one double per byte, zero elsewhere
        const int HB = (int)sizeof(cudaIpcMemHandle_t);
        std::vector<double> bytes((size_t)nranks * (HB + 1), 0.0);
        for (int i = 0; i < HB; ++i) bytes[(size_t)rank * (HB + 1) + i] = (double)((const unsigned char *)&h)[i];
        bytes[(size_t)rank * (HB + 1) + HB] = 1.0;  // "this rank has a mailbox"
        double *d = nullptr;
        SVB_CUDA(cudaMalloc((void **)&d, bytes.size() * 8));
        SVB_CUDA(cudaMemcpyAsync(d, bytes.data(), bytes.size() * 8, cudaMemcpyHostToDevice, C.stream));
        comm_allreduce_dev(d, (int64_t)bytes.size());
        SVB_CUDA(cudaMemcpyAsync(bytes.data(), d, bytes.size() * 8, cudaMemcpyDeviceToHost, C.stream));
        SVB_CUDA(cudaStreamSynchronize(C.stream));
        cudaFree(d);
        bool ok = true;
        for (int q = 0; q < nranks; ++q) ok = ok && bytes[(size_t)q * (HB + 1) + HB] == 1.0;
        for (int q = 0; q < nranks && ok; ++q) {
            if (q == rank) {
                g_p2p.peers_host[q] = g_p2p.mine;
                continue;
            }
            cudaIpcMemHandle_t hq;
            for (int i = 0; i < HB; ++i) ((unsigned char *)&hq)[i] = (unsigned char)bytes[(size_t)q * (HB + 1) + i];
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, hq, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                ok = false;
                break;
            }
            g_p2p.peers_host[q] = (Mailbox *)p;
        }
        // every rank must agree, otherwise some would wait on mailboxes nobody writes
        double agree = ok ? 1.0 : 0.0;
        double *da = nullptr;
        SVB_CUDA(cudaMalloc((void **)&da, 8));
        SVB_CUDA(cudaMemcpyAsync(da, &agree, 8, cudaMemcpyHostToDevice, C.stream));
        comm_allreduce_dev(da, 1);
        SVB_CUDA(cudaMemcpyAsync(&agree, da, 8, cudaMemcpyDeviceToHost, C.stream));
        SVB_CUDA(cudaStreamSynchronize(C.stream));
        cudaFree(da);
        if (agree != (double)nranks) {  // stay on NCCL; close what was opened
            p2p_teardown();
            return;
        }
        SVB_CUDA(cudaMalloc((void **)&g_p2p.peers_dev, sizeof(Mailbox *) * P2P_MAX_RANKS));
        SVB_CUDA(cudaMemcpyAsync(g_p2p.peers_dev, g_p2p.peers_host, sizeof(Mailbox *) * P2P_MAX_RANKS, cudaMemcpyHostToDevice, C.stream));
        SVB_CUDA(cudaStreamSynchronize(C.stream));
        SVB_CUDA(cudaMalloc((void **)&g_p2p.counter, 64));
        SVB_CUDA(cudaMemsetAsync(g_p2p.counter, 0, 64, C.stream));
        SVB_CUDA(cudaStreamSynchronize(C.stream));
        g_p2p.ready = true;
    } catch (const Error &) {
        p2p_teardown();  // stay on NCCL; nothing stays mapped
    }
}

// ---- one process, N GPUs (multi.cu): the workers live in one address space, so a peer's mailbox is reached through plain
// peer access (cudaDeviceEnablePeerAccess) — no IPC handles, no exchange through NCCL. Phase 1, every worker: allocate.
Mailbox *p2p_local_alloc() {
    g_p2p = P2PState();
    const char *env = getenv("SVB_P2P");
    if (env && atoi(env) == 0) return nullptr;
    Context &C = ctx();
    if (cudaMalloc((void **)&g_p2p.mine, sizeof(Mailbox)) != cudaSuccess) {
        cudaGetLastError();
        g_p2p.mine = nullptr;
        return nullptr;
    }
    cudaMemsetAsync(g_p2p.mine, 0, sizeof(Mailbox), C.stream);
    cudaStreamSynchronize(C.stream);
    return g_p2p.mine;
}

// Phase 2 (after every worker finished phase 1 and enabled peer access): the table of all mailboxes.
void p2p_local_connect(int nranks, int rank, Mailbox *const *all) {
    if (g_p2p.mine == nullptr) return;
    Context &C = ctx();
    g_p2p.ipc = false;
    g_p2p.nranks = nranks;
    g_p2p.rank = rank;
    for (int q = 0; q < nranks; ++q) g_p2p.peers_host[q] = all[q];
    try {
        SVB_CUDA(cudaMalloc((void **)&g_p2p.peers_dev, sizeof(Mailbox *) * P2P_MAX_RANKS));
        SVB_CUDA(cudaMemcpyAsync(g_p2p.peers_dev, g_p2p.peers_host, sizeof(Mailbox *) * P2P_MAX_RANKS, cudaMemcpyHostToDevice, C.stream));
        SVB_CUDA(cudaMalloc((void **)&g_p2p.counter, 64));
        SVB_CUDA(cudaMemsetAsync(g_p2p.counter, 0, 64, C.stream));
        SVB_CUDA(cudaStreamSynchronize(C.stream));
        g_p2p.ready = true;
    } catch (const Error &) {
        p2p_teardown();
    }
}

void p2p_teardown() {
    if (g_p2p.mine == nullptr) return;
    cudaDeviceSynchronize();
    if (g_p2p.ipc)
        for (int q = 0; q < g_p2p.nranks; ++q)
            if (q != g_p2p.rank && g_p2p.peers_host[q]) cudaIpcCloseMemHandle(g_p2p.peers_host[q]);
    if (g_p2p.peers_dev) cudaFree(g_p2p.peers_dev);
    if (g_p2p.counter) cudaFree(g_p2p.counter);
    cudaFree(g_p2p.mine);
    g_p2p = P2PState();
}

}  // namespace svb

// p2p.cuh — device-side pieces of the peer-memory exchange (see p2p.cu) so that producer / consumer kernels can
// fuse it: a producer stores its reduced values straight into every rank's mailbox and its last block
// publishes the epoch flag; a consumer waits on the local flags and sums the slots in rank order.
#pragma once
#include "svb_internal.h"

namespace svb {

constexpr int P2P_MAX_RANKS = 16;
constexpr int64_t P2P_CAP = 8192;  // doubles per slot (64 KB)

struct Mailbox {
    double slots[2][P2P_MAX_RANKS][P2P_CAP];
    unsigned long long flags[2][P2P_MAX_RANKS];
    int error;
};

// passed by value to kernels
struct P2PCtx {
    Mailbox *const *peers;  // device array [nranks] of mapped mailboxes (own entry = local pointer)
    Mailbox *mine;
    int nranks, rank;
    unsigned long long epoch;
    long long timeout_cycles;
    unsigned int *counter;  // device counter for last-block detection (self-resetting)
};

__device__ __forceinline__ void st_flag_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_flag_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// value `v` of element `idx` of this rank's contribution -> slot [parity][rank][idx] of every mailbox
__device__ __forceinline__ void p2p_store(const P2PCtx &c, int idx, double v) {
    const int par = (int)(c.epoch & 1ull);
    for (int q = 0; q < c.nranks; ++q) c.peers[q]->slots[par][c.rank][idx] = v;
}

// Called by EVERY thread of EVERY block of the producer after its p2p_store calls. The last block to arrive
// (all blocks fenced at system scope before taking a ticket) publishes the epoch flag in every mailbox.
__device__ __forceinline__ void p2p_publish_last_block(const P2PCtx &c, unsigned int nblocks) {
    __shared__ bool p2p_is_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(c.counter, 1u);
        p2p_is_last = (ticket == nblocks - 1);
    }
    __syncthreads();
    if (p2p_is_last) {
        __threadfence_system();
        const int par = (int)(c.epoch & 1ull);
        if ((int)threadIdx.x < c.nranks) st_flag_sys(&c.peers[threadIdx.x]->flags[par][c.rank], c.epoch);
        if (threadIdx.x == 0) *c.counter = 0u;
    }
}

// Called by every thread of a consumer block before p2p_sum: bounded wait for all ranks' flags of this epoch.
__device__ __forceinline__ void p2p_wait(const P2PCtx &c) {
    const int par = (int)(c.epoch & 1ull);
    if ((int)threadIdx.x < c.nranks) {
        const bool dead = *((volatile int *)&c.mine->error) != 0;
        const long long t0 = clock64();
        while (ld_flag_sys(&c.mine->flags[par][threadIdx.x]) < c.epoch) {
            if (dead || clock64() - t0 > c.timeout_cycles) {
                c.mine->error = 1;
                break;
            }
        }
    }
    __syncthreads();
}

// rank-ordered sum of element idx (identical bits on every rank); L1 bypassed (peers write through NVLink into L2)
__device__ __forceinline__ double p2p_sum(const P2PCtx &c, int idx) {
    const int par = (int)(c.epoch & 1ull);
    double s = 0.0;
    for (int q = 0; q < c.nranks; ++q) s += __ldcg(&c.mine->slots[par][q][idx]);
    return s;
}

// host side
bool p2p_ready();
bool p2p_next_ctx(int64_t n, P2PCtx *out);  // reserves the next epoch; false when the path is unavailable or n too large

}  // namespace svb

// core.cu — library context, error reporting, per-class device timers, device prefix scan.
#define SVB_NO_ALLOC_MACROS
#include "svb_internal.h"

#include <cstring>
#include <mutex>
#include <unordered_map>

namespace svb {

static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }

static thread_local Context *tl_ctx = nullptr;
Context &ctx() {
    static Context primary;
    return tl_ctx ? *tl_ctx : primary;
}
void set_thread_context(Context *c) { tl_ctx = c; }

void require_init() {
    if (!ctx().initialised)
        throw Error(SVB_ECUDA, "svb_init has not been called (or no CUDA device): there is no CPU fallback");
}

void count_launch(int n) { ctx().launches += n; }

// Device allocator. Every allocation of the library is made, used and freed in the order of ONE stream, so a freed
// block can be handed to the next request without any driver call or synchronisation: blocks are cached in size
// classes (8 per power of two, <= 12.5 % slack) on top of cudaMalloc and never returned to the driver before
// svb_shutdown (or an out-of-memory retry). Round 1 used the stream-ordered pool (cudaMallocAsync, release threshold
// unlimited): no cudaFree stall any more, but at C3 scale the pool re-maps physical memory whenever a multi-GB
// request does not fit a free virtual range, and identical operator builds took anywhere from 0.12 s to 1.5 s
// (profiles/r02_counts_operator.md). SVB_ALLOC=pool selects the old behaviour.
namespace {
struct BlockCache {
    std::mutex mu;
    std::unordered_map<size_t, std::vector<void *>> free_blocks;  // size class -> cached blocks
    std::unordered_map<void *, size_t> live;                      // block -> size class
    size_t cached_bytes = 0;
    bool use_pool = getenv("SVB_ALLOC") != nullptr && std::string(getenv("SVB_ALLOC")) == "pool";
};
BlockCache &cache() {  // one cache per context: blocks belong to that context's device and stream
    Context &C = ctx();
    if (!C.alloc_cache) C.alloc_cache = new BlockCache();
    return *static_cast<BlockCache *>(C.alloc_cache);
}
size_t size_class(size_t bytes) {
    if (bytes <= 512) return 512;
    int lg = 63 - __builtin_clzll((unsigned long long)bytes);
    const size_t step = (size_t)1 << std::max(lg - 3, 9);
    return (bytes + step - 1) / step * step;
}
void release_cached_blocks() {
    BlockCache &B = cache();
    for (auto &kv : B.free_blocks)
        for (void *p : kv.second) cudaFree(p);
    B.free_blocks.clear();
    B.cached_bytes = 0;
}
}  // namespace

cudaError_t dev_malloc(void **p, size_t bytes) {
    Context &C = ctx();
    BlockCache &B = cache();
    if (bytes == 0) bytes = 8;
    if (B.use_pool) {
        if (C.initialised && C.stream && C.pool_ok) return cudaMallocAsync(p, bytes, C.stream);
        return cudaMalloc(p, bytes);
    }
    const size_t cls = size_class(bytes);
    std::lock_guard<std::mutex> lock(B.mu);
    auto it = B.free_blocks.find(cls);
    if (it != B.free_blocks.end() && !it->second.empty()) {
        *p = it->second.back();
        it->second.pop_back();
        B.cached_bytes -= cls;
        B.live[*p] = cls;
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, cls);
    if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and retry once
        cudaGetLastError();
        if (C.stream) cudaStreamSynchronize(C.stream);
        release_cached_blocks();
        e = cudaMalloc(p, cls);
    }
    if (e == cudaSuccess) B.live[*p] = cls;
    return e;
}

cudaError_t dev_free(void *p) {
    Context &C = ctx();
    BlockCache &B = cache();
    if (!p) return cudaSuccess;
    if (B.use_pool) {
        if (C.initialised && C.stream && C.pool_ok) return cudaFreeAsync(p, C.stream);
        return cudaFree(p);  // valid for pool memory too (synchronises)
    }
    std::lock_guard<std::mutex> lock(B.mu);
    auto it = B.live.find(p);
    if (it == B.live.end()) return cudaFree(p);
    const size_t cls = it->second;
    B.live.erase(it);
    B.free_blocks[cls].push_back(p);  // reusable at once: the next user is ordered after this one on the library stream
    B.cached_bytes += cls;
    return cudaSuccess;
}

void dev_release_cache() {
    std::lock_guard<std::mutex> lock(cache().mu);
    release_cached_blocks();
}

KTimer::KTimer(int c, double algorithmic_bytes, int nlaunch) : cls(c), bytes(algorithmic_bytes) {
    Context &C = ctx();
    C.launches += nlaunch;
    if (C.profile) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, C.stream);
    }
    if (c >= 0 && c < SVB_K_NCLASS) {
        C.prof_launches[c] += nlaunch;
        C.prof_bytes[c] += algorithmic_bytes;
    }
}

KTimer::~KTimer() {
    Context &C = ctx();
    if (e0) {
        cudaEventRecord(e1, C.stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (cls >= 0 && cls < SVB_K_NCLASS) C.prof_ms[cls] += ms;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
}

// ---------------------------------------------------------------------------------------------
// exclusive scan (int64, in place). Three passes: per-block scan + block totals, scan of the
// totals (recursive), add offsets. 1024 threads x 4 items per block.
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_T = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_BLOCK = SCAN_T * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_T) scan_block_kernel(int64_t *data, int64_t n, int64_t *block_sums) {
    __shared__ int64_t warp_tot[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? data[base + i] : 0;
        tsum += v[i];
    }
    // inclusive scan of tsum across the block
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int64_t x = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int64_t w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    int64_t excl = x - tsum + (wid > 0 ? warp_tot[wid - 1] : 0);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) data[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == SCAN_T - 1 && block_sums) block_sums[blockIdx.x] = excl;
}

__global__ void scan_add_kernel(int64_t *data, int64_t n, const int64_t *block_offs) {
    const int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const int64_t off = block_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t idx = i + (int64_t)k * SCAN_T;
        if (idx < n) data[idx] += off;
    }
}

void exclusive_scan_i64(int64_t *d, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    const int64_t nblk = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (nblk == 1) {
        scan_block_kernel<<<1, SCAN_T, 0, st>>>(d, n, nullptr);
        count_launch();
        SVB_CUDA(cudaGetLastError());
        return;
    }
    DevBuf<int64_t> sums((size_t)nblk);
    scan_block_kernel<<<(unsigned)nblk, SCAN_T, 0, st>>>(d, n, sums.p);
    count_launch();
    SVB_CUDA(cudaGetLastError());
    exclusive_scan_i64(sums.p, nblk, st);
    scan_add_kernel<<<(unsigned)nblk, SCAN_T, 0, st>>>(d, n, sums.p);
    count_launch();
    SVB_CUDA(cudaGetLastError());
    SVB_CUDA(cudaStreamSynchronize(st));  // sums freed on return
}

}  // namespace svb

using namespace svb;

extern "C" {

const char *svb_last_error(void) { return g_last_error.c_str(); }
const char *svb_version(void) { return "severo_b200 0.1 (sm_100a)"; }

int svb_init(int device) {
    SVB_API_BEGIN
    Context &C = ctx();
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw Error(SVB_ECUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                   "); libsevero_b200 has no CPU fallback");
    SVB_CHECK(device >= 0 && device < ndev, SVB_EARG, "svb_init: device index out of range");
    SVB_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    SVB_CUDA(cudaGetDeviceProperties(&p, device));
    C.device = device;
    C.sm_count = p.multiProcessorCount;
    C.smem_optin = p.sharedMemPerBlockOptin;
    if (!C.stream) {
        SVB_CUDA(cudaStreamCreateWithFlags(&C.stream, cudaStreamNonBlocking));
        C.own_stream = true;
    }
    {
        int pools = 0;
        cudaDeviceGetAttribute(&pools, cudaDevAttrMemoryPoolsSupported, device);
        cudaMemPool_t pool;
        if (pools && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            C.pool_ok = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr) == cudaSuccess;
        }
        cudaGetLastError();
    }
    C.initialised = true;
    SVB_API_END
}

int svb_shutdown(void) {
    SVB_API_BEGIN
    Context &C = ctx();
    if (C.initialised) {
        cudaDeviceSynchronize();
        dev_release_cache();
        if (C.own_stream && C.stream) cudaStreamDestroy(C.stream);
        C.stream = nullptr;
        C.own_stream = false;
        C.initialised = false;
    }
    SVB_API_END
}

int svb_set_stream(void *s) {
    SVB_API_BEGIN
    require_init();
    Context &C = ctx();
    SVB_CUDA(cudaStreamSynchronize(C.stream));  // cached blocks freed under the old stream are idle from here on
    if (C.own_stream && C.stream) cudaStreamDestroy(C.stream);
    if (s) {
        C.stream = (cudaStream_t)s;
        C.own_stream = false;
    } else {
        SVB_CUDA(cudaStreamCreateWithFlags(&C.stream, cudaStreamNonBlocking));
        C.own_stream = true;
    }
    SVB_API_END
}

int svb_synchronize(void) {
    SVB_API_BEGIN
    require_init();
    SVB_CUDA(cudaStreamSynchronize(ctx().stream));
    SVB_API_END
}

int svb_device_info(int *sm_count, int64_t *total_mem, int *cc_major, int *cc_minor) {
    SVB_API_BEGIN
    require_init();
    cudaDeviceProp p;
    SVB_CUDA(cudaGetDeviceProperties(&p, ctx().device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (total_mem) *total_mem = (int64_t)p.totalGlobalMem;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    SVB_API_END
}

int svb_profile_enable(int on) {
    ctx().profile = on != 0;
    return SVB_OK;
}

int svb_profile_reset(void) {
    Context &C = ctx();
    for (int i = 0; i < SVB_K_NCLASS; ++i) {
        C.prof_ms[i] = 0;
        C.prof_launches[i] = 0;
        C.prof_bytes[i] = 0;
    }
    return SVB_OK;
}

int svb_profile_get(double *ms, int64_t *launches, double *bytes) {
    Context &C = ctx();
    for (int i = 0; i < SVB_K_NCLASS; ++i) {
        if (ms) ms[i] = C.prof_ms[i];
        if (launches) launches[i] = C.prof_launches[i];
        if (bytes) bytes[i] = C.prof_bytes[i];
    }
    return SVB_OK;
}

int64_t svb_launch_count(void) { return ctx().launches; }
int svb_launch_count_reset(void) {
    ctx().launches = 0;
    return SVB_OK;
}

}  // extern "C"

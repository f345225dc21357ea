// comm.cpp — the one exchange step of the path: on N GPUs the cells are sharded, S'*w yields a
// gene-length partial per rank, and the Lanczos recurrences need its sum (plus the reorthogonalisation
// coefficients and norms of the cell-sharded basis). One process per GPU; NCCL over NVLink.
// libnccl is resolved at run time (dlopen) so the library also loads where NCCL is absent and a
// single-GPU caller never needs it.
#include "svb_internal.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

namespace svb {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi &nccl() {
    static NcclApi api;
    if (api.handle) return api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) throw Error(SVB_ENCCL, std::string("cannot load libnccl: ") + dlerror());
    auto sym = [&](const char *s) {
        void *p = dlsym(api.handle, s);
        if (!p) throw Error(SVB_ENCCL, std::string("libnccl lacks symbol ") + s);
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    return api;
}

#define SVB_NCCL(expr)                                                                          \
    do {                                                                                        \
        ncclResult_t _r = (expr);                                                               \
        if (_r != ncclSuccess)                                                                  \
            throw svb::Error(SVB_ENCCL, std::string("NCCL error: ") + nccl().GetErrorString(_r)); \
    } while (0)

void comm_allreduce_dev(double *dbuf, int64_t n) {
    Context &C = ctx();
    if (C.nranks <= 1 || n <= 0) return;
    if (p2p_allreduce(dbuf, n)) return;  // small message: one-shot kernel over NVLink peer memory
    KTimer kt(SVB_K_COMM, 8.0 * n, 1);
    SVB_NCCL(nccl().AllReduce(dbuf, dbuf, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)C.nccl_comm, C.stream));
}

// one process, N GPUs (multi.cu): the communicators of all devices in one call, no unique id to pass around
void comm_init_all(int ndev, const int *devs, void **comms_out) {
    std::vector<ncclComm_t> comms((size_t)ndev);
    int cur = 0;
    cudaGetDevice(&cur);
    SVB_NCCL(nccl().CommInitAll(comms.data(), ndev, devs));
    cudaSetDevice(cur);
    for (int i = 0; i < ndev; ++i) comms_out[i] = comms[(size_t)i];
}

void comm_destroy_one(void *comm) {
    if (comm) nccl().CommDestroy((ncclComm_t)comm);
}

}  // namespace svb

using namespace svb;

extern "C" {

int svb_comm_unique_id(unsigned char id[128]) {
    SVB_API_BEGIN
    SVB_CHECK(id, SVB_EARG, "svb_comm_unique_id: null pointer");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    SVB_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, &u, 128);
    SVB_API_END
}

int svb_comm_init(int nranks, int rank, const unsigned char id[128]) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, SVB_EARG, "svb_comm_init: bad rank / nranks");
    Context &C = ctx();
    SVB_CHECK(C.nccl_comm == nullptr, SVB_EARG, "svb_comm_init: communicator already initialised");
    if (nranks == 1) {
        C.nranks = 1;
        C.rank = 0;
        return SVB_OK;
    }
    SVB_CHECK(id, SVB_EARG, "svb_comm_init: null id");
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t comm;
    SVB_NCCL(nccl().CommInitRank(&comm, nranks, u, rank));
    C.nccl_comm = comm;
    C.nranks = nranks;
    C.rank = rank;
    p2p_setup(nranks, rank);
    SVB_API_END
}

int svb_comm_destroy(void) {
    SVB_API_BEGIN
    Context &C = ctx();
    p2p_teardown();
    if (C.nccl_comm) {
        cudaStreamSynchronize(C.stream);
        nccl().CommDestroy((ncclComm_t)C.nccl_comm);
        C.nccl_comm = nullptr;
    }
    C.nranks = 1;
    C.rank = 0;
    SVB_API_END
}

int svb_comm_info(int *nranks, int *rank) {
    if (nranks) *nranks = ctx().nranks;
    if (rank) *rank = ctx().rank;
    return SVB_OK;
}

int svb_comm_allreduce_f64(double *host_buf, int64_t n) {
    SVB_API_BEGIN
    require_init();
    SVB_CHECK(host_buf || n == 0, SVB_EARG, "svb_comm_allreduce_f64: null buffer");
    Context &C = ctx();
    if (C.nranks <= 1 || n <= 0) return SVB_OK;
    DevBuf<double> d((size_t)n);
    SVB_CUDA(cudaMemcpyAsync(d.p, host_buf, (size_t)n * 8, cudaMemcpyHostToDevice, C.stream));
    comm_allreduce_dev(d.p, n);
    SVB_CUDA(cudaMemcpyAsync(host_buf, d.p, (size_t)n * 8, cudaMemcpyDeviceToHost, C.stream));
    SVB_CUDA(cudaStreamSynchronize(C.stream));
    SVB_API_END
}

}  // extern "C"

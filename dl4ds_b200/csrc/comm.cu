// dl4ds_comm_*: the data-parallel exchange of the hot path behind the C ABI -- what the
// reference delegates to Horovod (hvd.DistributedOptimizer / DistributedGradientTape / broadcast_variables:
// training/supervised.py:363-369, training/cgan.py:608-637) as NCCL collectives over NVLink / NVSwitch on the flat
// gradient arena.  One communicator per process (one process per GPU).  Every call takes the caller's stream, is
// asynchronous, and can be captured into a CUDA graph (the optimizer step is ONE graph: forward, backward,
// all-reduce, Adam).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the process already loaded -- PyTorch bundles one -- or
// the system library), so libdl4ds_b200.so carries no link-time dependency on it and single-GPU use never touches it.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace dl4ds {

namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclFloat = 7, kNcclSum = 0, kNcclChar = 0 };

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Nccl g_nccl;
ncclComm_t g_comm = nullptr;
int g_nranks = 0, g_rank = -1;

int load_nccl() {
    if (g_nccl.handle) return DL4DS_OK;
    const char* names[] = {getenv("DL4DS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !n[0]) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("dl4ds_comm: cannot load libnccl.so.2 (%s); set DL4DS_NCCL_LIB", dlerror());
        return DL4DS_E_NCCL;
    }
#define NCCL_SYM(field, sym)                                                                     \
    do {                                                                                         \
        *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, sym);                                \
        if (!g_nccl.field) {                                                                     \
            set_error("dl4ds_comm: symbol %s missing in libnccl", sym);                          \
            return DL4DS_E_NCCL;                                                                 \
        }                                                                                        \
    } while (0)
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    NCCL_SYM(CommInitRank, "ncclCommInitRank");
    NCCL_SYM(CommDestroy, "ncclCommDestroy");
    NCCL_SYM(AllReduce, "ncclAllReduce");
    NCCL_SYM(Broadcast, "ncclBroadcast");
    NCCL_SYM(GetVersion, "ncclGetVersion");
    NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
    g_nccl.handle = h;
    return DL4DS_OK;
}

int nccl_check(ncclResult_t r, const char* what) {
    if (r == 0) return DL4DS_OK;
    set_error("%s: NCCL error %d (%s)", what, (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return DL4DS_E_NCCL;
}

}  // namespace

}  // namespace dl4ds

using namespace dl4ds;

extern "C" {

int dl4ds_comm_unique_id_bytes(void) { return 128; }

int dl4ds_comm_nccl_version(void) {
    if (load_nccl() != DL4DS_OK) return -1;
    int v = 0;
    if (g_nccl.GetVersion(&v) != 0) return -1;
    return v;
}

int dl4ds_comm_get_unique_id(void* id_out_128) {
    DL4DS_REQUIRE(id_out_128 != nullptr, DL4DS_E_BADARG, "comm_get_unique_id: null pointer");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    rc = nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
    if (rc) return rc;
    memcpy(id_out_128, &id, sizeof(id));
    return DL4DS_OK;
}

int dl4ds_comm_init_rank(const void* id_128, int nranks, int rank) {
    DL4DS_REQUIRE(id_128 != nullptr && nranks >= 1 && rank >= 0 && rank < nranks, DL4DS_E_BADARG, "comm_init_rank: bad argument");
    DL4DS_REQUIRE(g_comm == nullptr, DL4DS_E_BADARG, "comm_init_rank: a communicator already exists (dl4ds_comm_destroy first)");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, id_128, sizeof(id));
    rc = nccl_check(g_nccl.CommInitRank(&g_comm, nranks, id, rank), "ncclCommInitRank");
    if (rc) { g_comm = nullptr; return rc; }
    g_nranks = nranks;
    g_rank = rank;
    return DL4DS_OK;
}

int dl4ds_comm_size(void) { return g_comm ? g_nranks : 0; }
int dl4ds_comm_rank(void) { return g_comm ? g_rank : -1; }

int dl4ds_comm_allreduce_sum(float* buf, int64_t n, void* stream) {
    DL4DS_REQUIRE(g_comm != nullptr, DL4DS_E_BADARG, "comm_allreduce_sum: no communicator (dl4ds_comm_init_rank)");
    DL4DS_REQUIRE(buf != nullptr && n >= 0, DL4DS_E_BADARG, "comm_allreduce_sum: bad argument");
    if (n == 0) return DL4DS_OK;
    return nccl_check(g_nccl.AllReduce(buf, buf, (size_t)n, kNcclFloat, kNcclSum, g_comm, reinterpret_cast<cudaStream_t>(stream)),
                      "ncclAllReduce");
}

int dl4ds_comm_broadcast(void* buf, int64_t nbytes, int root, void* stream) {
    DL4DS_REQUIRE(g_comm != nullptr, DL4DS_E_BADARG, "comm_broadcast: no communicator (dl4ds_comm_init_rank)");
    DL4DS_REQUIRE(buf != nullptr && nbytes >= 0 && root >= 0 && root < g_nranks, DL4DS_E_BADARG, "comm_broadcast: bad argument");
    if (nbytes == 0) return DL4DS_OK;
    return nccl_check(g_nccl.Broadcast(buf, buf, (size_t)nbytes, kNcclChar, root, g_comm, reinterpret_cast<cudaStream_t>(stream)),
                      "ncclBroadcast");
}

int dl4ds_comm_destroy(void) {
    if (!g_comm) return DL4DS_OK;
    const int rc = nccl_check(g_nccl.CommDestroy(g_comm), "ncclCommDestroy");
    g_comm = nullptr;
    g_nranks = 0;
    g_rank = -1;
    return rc;
}

}  // extern "C"

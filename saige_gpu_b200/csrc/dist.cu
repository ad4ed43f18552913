// Multi-GPU plumbing: one rank per GPU, markers sharded block-cyclically, ONE sum-allreduce per (multi-)product.
// Replaces MPI_Barrier + host-staged MPI_Allreduce(MPI_FLOAT) (FG.cpp:1614-1620, 1647-1653) with an in-place
// device ncclAllReduce(ncclDouble) over NVLink 5 / NVSwitch.  NCCL is resolved with dlopen at first use so that a
// single-GPU process needs no NCCL at all (and a host process that already loaded torch's libnccl shares it).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <string>
#include "sgb_internal.h"

struct nccl_api {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

struct sgb_dist {
    ncclComm_t comm = nullptr;
};

static nccl_api g_nccl;

static int load_nccl(std::string &err)
{
    if (g_nccl.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { err = std::string("cannot load NCCL (libnccl.so.2): ") + dlerror(); return 1; }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(lib, "ncclAllReduce");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast))dlsym(lib, "ncclBroadcast");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.Broadcast || !g_nccl.GetErrorString) {
        err = "libnccl.so.2 lacks a required symbol";
        return 1;
    }
    g_nccl.lib = lib;
    return 0;
}

int sgb_dist_unique_id(void *id128, std::string &err)
{
    static_assert(sizeof(ncclUniqueId) == SGB_NCCL_ID_BYTES, "ncclUniqueId size");
    if (load_nccl(err)) return 1;
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return 1; }
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int sgb_dist_init(sgb_ctx *h, int rank, int world, const void *id128)
{
    std::string err;
    if (load_nccl(err)) return sgb_fail(h, "%s", err.c_str());
    if (rank < 0 || rank >= world) return sgb_fail(h, "bad rank %d of %d", rank, world);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    h->dist = new sgb_dist();
    CUDA_OK(h, cudaSetDevice(h->device));
    ncclResult_t r = g_nccl.CommInitRank(&h->dist->comm, world, id, rank);
    if (r != ncclSuccess) return sgb_fail(h, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    h->rank = rank; h->world = world;
    return 0;
}

void sgb_dist_destroy(sgb_ctx *h)
{
    if (!h->dist) return;
    if (h->dist->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->dist->comm);
    delete h->dist;
    h->dist = nullptr;
}

int sgb_allreduce_sum(sgb_ctx *h, double *d, int64_t n)
{
    if (h->world <= 1) return 0;
    if (!h->dist || !h->dist->comm) return sgb_fail(h, "allreduce requested but NCCL communicator is not initialised");
    SGB_RANGE("nccl_allreduce");
    ncclResult_t r = g_nccl.AllReduce(d, d, (size_t)n, ncclDouble, ncclSum, h->dist->comm, h->stream);
    if (r != ncclSuccess) return sgb_fail(h, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
    h->cnt.n_allreduce++;
    return 0;
}

int sgb_allreduce_sum_i32(sgb_ctx *h, int32_t *d, int64_t n)
{
    if (h->world <= 1) return 0;
    if (!h->dist || !h->dist->comm) return sgb_fail(h, "allreduce requested but NCCL communicator is not initialised");
    ncclResult_t r = g_nccl.AllReduce(d, d, (size_t)n, ncclInt32, ncclSum, h->dist->comm, h->stream);
    if (r != ncclSuccess) return sgb_fail(h, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
    h->cnt.n_allreduce++;
    return 0;
}

// in-place broadcast of a device buffer from `root` (dense-GRM build: every rank needs every marker shard)
int sgb_broadcast_bytes(sgb_ctx *h, void *d, size_t bytes, int root)
{
    if (h->world <= 1) return 0;
    if (!h->dist || !h->dist->comm) return sgb_fail(h, "broadcast requested but NCCL communicator is not initialised");
    ncclResult_t r = g_nccl.Broadcast(d, d, bytes, ncclUint8, root, h->dist->comm, h->stream);
    if (r != ncclSuccess) return sgb_fail(h, "ncclBroadcast: %s", g_nccl.GetErrorString(r));
    return 0;
}

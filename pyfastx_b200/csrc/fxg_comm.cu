// fxg_comm.cu -- the one small exchange of the multi-GPU index build: an in-stream
// ncclAllGather of 128-byte fxg_shard_info structs over NVLink / NVSwitch.  NCCL is bound at run time with
// dlopen("libnccl.so.2") -- the copy already mapped into the process (e.g. torch's bundled one) if there is one --
// so libfxg.so has no link-time dependency on a particular NCCL build and single-GPU use never loads it.
#include "fxg_common.cuh"
#include <dlfcn.h>
#include <string.h>

namespace {

struct NcclUniqueId { char internal[FXG_COMM_ID_BYTES]; };   // == ncclUniqueId (NCCL_UNIQUE_ID_BYTES 128)
typedef void *NcclComm;
typedef int (*fn_GetUniqueId)(NcclUniqueId *);
typedef int (*fn_CommInitRank)(NcclComm *, int, NcclUniqueId, int);
typedef int (*fn_CommDestroy)(NcclComm);
typedef int (*fn_AllGather)(const void *, void *, size_t, int /*ncclDataType_t*/, NcclComm, cudaStream_t);
typedef const char *(*fn_GetErrorString)(int);
typedef int (*fn_GetVersion)(int *);

struct NcclApi {
    void *handle = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_AllGather AllGather = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
    fn_GetVersion GetVersion = nullptr;
};

std::mutex g_mu;
NcclApi g_api;

int load_nccl() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_api.handle) return FXG_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    void *h = nullptr;
    // RTLD_NOLOAD first: reuse the NCCL the process already runs (two NCCL copies in one process must be avoided)
    for (int i = 0; names[i] && !h; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    for (int i = 0; names[i] && !h; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fxg_set_error("cannot load libnccl.so.2: %s", dlerror()); return FXG_ENODEV; }
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = (fn_GetUniqueId)dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (fn_CommInitRank)dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (fn_CommDestroy)dlsym(h, "ncclCommDestroy");
    a.AllGather = (fn_AllGather)dlsym(h, "ncclAllGather");
    a.GetErrorString = (fn_GetErrorString)dlsym(h, "ncclGetErrorString");
    a.GetVersion = (fn_GetVersion)dlsym(h, "ncclGetVersion");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather) {
        fxg_set_error("libnccl.so.2 lacks a required symbol");
        return FXG_ENODEV;
    }
    g_api = a;
    return FXG_OK;
}

int nccl_fail(const char *what, int code) {
    fxg_set_error("%s failed: %s", what, g_api.GetErrorString ? g_api.GetErrorString(code) : "NCCL error");
    return FXG_ECUDA;
}

}  // namespace

struct fxg_comm {
    NcclComm comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
};

extern "C" int fxg_comm_unique_id(void *id_out) {
    FXG_CHECK_ARG(id_out, "id_out == NULL");
    int rc = load_nccl();
    if (rc) return rc;
    NcclUniqueId id;
    const int e = g_api.GetUniqueId(&id);
    if (e) return nccl_fail("ncclGetUniqueId", e);
    memcpy(id_out, &id, sizeof(id));
    return FXG_OK;
}

extern "C" int fxg_comm_create(fxg_ctx *ctx, const void *id, int nranks, int rank, fxg_comm **out) {
    FXG_CHECK_ARG(ctx && id && out && nranks >= 1 && rank >= 0 && rank < nranks, "bad arguments");
    *out = nullptr;
    FXG_LOCK(ctx);
    int rc = load_nccl();
    if (rc) return rc;
    FXG_CUDA(cudaSetDevice(ctx->device));
    NcclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    fxg_comm *c = new fxg_comm();
    c->nranks = nranks; c->rank = rank; c->device = ctx->device;
    const int e = g_api.CommInitRank(&c->comm, nranks, uid, rank);
    if (e) { delete c; return nccl_fail("ncclCommInitRank", e); }
    *out = c;
    return FXG_OK;
}

extern "C" int fxg_comm_nranks(const fxg_comm *c) { return c ? c->nranks : 1; }
extern "C" int fxg_comm_rank(const fxg_comm *c) { return c ? c->rank : 0; }

extern "C" void fxg_comm_destroy(fxg_comm *c) {
    if (!c) return;
    if (c->comm && g_api.CommDestroy) { cudaSetDevice(c->device); g_api.CommDestroy(c->comm); }
    delete c;
}

// In-stream all-gather of `bytes_per_rank` bytes from every rank (rank order) on the context's stream.
extern "C" int fxg_shard_exchange(fxg_ctx *ctx, fxg_comm *comm, const void *d_send, void *d_recv, int64_t bytes_per_rank) {
    FXG_CHECK_ARG(ctx && d_send && d_recv && bytes_per_rank > 0, "bad arguments");
    FXG_LOCK(ctx);
    FXG_CUDA(cudaSetDevice(ctx->device));
    if (!comm || comm->nranks == 1) {
        if (d_send != d_recv)
            FXG_CUDA(cudaMemcpyAsync(d_recv, d_send, (size_t)bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return FXG_OK;
    }
    const int e = g_api.AllGather(d_send, d_recv, (size_t)bytes_per_rank, /*ncclInt8*/ 0, comm->comm, ctx->stream);
    if (e) return nccl_fail("ncclAllGather", e);
    ctx->collectives += 1;
    return FXG_OK;
}

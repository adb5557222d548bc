// fxg_comm.cu -- the one small exchange of the multi-GPU index build: an in-stream
// ncclAllGather of 128-byte fxg_shard_info structs over NVLink / NVSwitch.  NCCL is bound at run time with
// dlopen("libnccl.so.2") -- the copy already mapped into the process (e.g. torch's bundled one) if there is one --
// so libfxg.so has no link-time dependency on a particular NCCL build and single-GPU use never loads it.
#include "fxg_common.cuh"
#include <dlfcn.h>
#include <string.h>
#include <stdlib.h>
#include <vector>

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

// ---- peer-memory mailboxes (NVLink / NVSwitch P2P stores): the exchange of the sharded scan without NCCL ----
// Every rank owns a mailbox in its HBM: 2 (epoch parity) x nranks slots of FXG_MBOX_SLOT bytes, followed by
// 2 x nranks 64-bit flags.  In an exchange, rank i stores its block into slot [epoch & 1][i] of EVERY rank's mailbox
// (its own included), fences system-wide and then stores the epoch number into flag [epoch & 1][i] there; it then
// waits until all nranks flags of its own mailbox show the epoch and copies the slots out.  One kernel of one warp per
// rank, no host involvement, no collective library on the path.  A slot is reused two exchanges later: a rank can
// only reach exchange e + 2 after every peer finished reading exchange e (it has seen their e + 1 flags, which they
// set after copying e out, in stream order).
constexpr int FXG_MBOX_SLOT = 256;          // bytes per rank and exchange
constexpr int FXG_MBOX_MAXR = 64;

struct MboxPtrs { uint8_t *p[FXG_MBOX_MAXR]; };

__global__ void __launch_bounds__(256) mbox_exchange_kernel(MboxPtrs peers, int nranks, int rank, unsigned long long epoch,
                                                            const uint8_t *__restrict__ send, uint8_t *__restrict__ recv,
                                                            int bytes, int *__restrict__ status) {
    const int par = (int)(epoch & 1ull);
    const int pieces = bytes / 16;                                      // 16-byte pieces per block
    const size_t flag_base = (size_t)2 * nranks * FXG_MBOX_SLOT;
    // 1. my block -> every rank's mailbox
    for (int i = threadIdx.x; i < nranks * pieces; i += blockDim.x) {
        const int r = i / pieces, k = i % pieces;
        const uint4 v = reinterpret_cast<const uint4 *>(send)[k];
        uint4 *dst = reinterpret_cast<uint4 *>(peers.p[r] + ((size_t)par * nranks + rank) * FXG_MBOX_SLOT) + k;
        asm volatile("st.global.relaxed.sys.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: my flag in every rank's mailbox
    if ((int)threadIdx.x < nranks) {
        unsigned long long *f = reinterpret_cast<unsigned long long *>(peers.p[threadIdx.x] + flag_base) + (size_t)par * nranks + rank;
        asm volatile("st.global.release.sys.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    }
    // 3. wait for every rank's flag in MY mailbox (bounded: a dead peer must not hang this GPU)
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    if ((int)threadIdx.x < nranks) {
        const unsigned long long *f = reinterpret_cast<const unsigned long long *>(peers.p[rank] + flag_base) + (size_t)par * nranks + threadIdx.x;
        const long long t0 = clock64();
        unsigned long long v;
        for (;;) {
            asm volatile("ld.global.acquire.sys.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= epoch) break;
            if (clock64() - t0 > 20000000000ll) { s_fail = 1; break; }            // ~10 s at 2 GHz
        }
    }
    __syncthreads();
    if (s_fail) { if (threadIdx.x == 0) *status = 1; return; }
    // 4. gathered blocks -> recv (rank order)
    for (int i = threadIdx.x; i < nranks * pieces; i += blockDim.x) {
        const int r = i / pieces, k = i % pieces;
        const uint4 *src = reinterpret_cast<const uint4 *>(peers.p[rank] + ((size_t)par * nranks + r) * FXG_MBOX_SLOT) + k;
        uint4 v;
        asm volatile("ld.global.relaxed.sys.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
        reinterpret_cast<uint4 *>(recv + (size_t)r * bytes)[k] = v;
    }
}

struct fxg_comm {
    NcclComm comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    // peer-memory path
    bool p2p = false;
    uint8_t *mbox = nullptr;                 // this rank's mailbox (cudaMalloc)
    MboxPtrs peers = {};                     // every rank's mailbox as seen from this device (IPC mappings; [rank] = mbox)
    unsigned long long epoch = 0;
    int *h_status = nullptr;                 // mapped pinned host word: the kernel reports a timed-out wait here
    int *d_status = nullptr;                 // its device alias
};

// Sets up the mailboxes: allocate, exchange the IPC handles with ONE ncclAllGather at creation time, map the peers'.
// Any failure (IPC not permitted, no peer access) leaves p2p off and the exchange on ncclAllGather.
static void setup_p2p(fxg_comm *c, cudaStream_t stream) {
    const char *e = getenv("FXG_COMM");
    if (e && !strcmp(e, "nccl")) return;
    if (c->nranks > FXG_MBOX_MAXR) return;
    const size_t bytes = (size_t)2 * c->nranks * FXG_MBOX_SLOT + (size_t)2 * c->nranks * 8;
    cudaIpcMemHandle_t mine;
    uint8_t *d_h = nullptr;
    bool ok = cudaMalloc((void **)&c->mbox, bytes) == cudaSuccess && cudaMemset(c->mbox, 0, bytes) == cudaSuccess &&
              cudaHostAlloc((void **)&c->h_status, sizeof(int), cudaHostAllocMapped) == cudaSuccess &&
              cudaHostGetDevicePointer((void **)&c->d_status, c->h_status, 0) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine, c->mbox) == cudaSuccess &&
              cudaMalloc((void **)&d_h, (size_t)(c->nranks + 1) * sizeof(mine)) == cudaSuccess;
    std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
    // every rank takes part in the gather, also one whose local set-up failed (it sends zeros and everybody falls back)
    int local_ok = ok ? 1 : 0;
    if (ok) *c->h_status = 0;
    if (!ok) { cudaGetLastError(); memset(&mine, 0, sizeof(mine)); if (!d_h) cudaMalloc((void **)&d_h, (size_t)(c->nranks + 1) * sizeof(mine)); }
    if (!d_h) return;                                                     // cannot even gather: NCCL path (peers time out on IPC open? no: see below)
    uint8_t *d_mine = d_h + (size_t)c->nranks * sizeof(mine);
    cudaMemcpyAsync(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice, stream);
    const int ge = g_api.AllGather(d_mine, d_h, sizeof(mine), /*ncclInt8*/ 0, c->comm, stream);
    cudaMemcpyAsync(all.data(), d_h, (size_t)c->nranks * sizeof(mine), cudaMemcpyDeviceToHost, stream);
    const bool gathered = ge == 0 && cudaStreamSynchronize(stream) == cudaSuccess;
    cudaFree(d_h);
    if (!gathered || !local_ok) { cudaGetLastError(); return; }
    cudaIpcMemHandle_t zero;
    memset(&zero, 0, sizeof(zero));
    for (int r = 0; r < c->nranks; ++r) {
        if (!memcmp(&all[(size_t)r], &zero, sizeof(zero))) return;      // a peer could not set up: everybody stays on NCCL
    }
    for (int r = 0; r < c->nranks; ++r) {
        if (r == c->rank) { c->peers.p[r] = c->mbox; continue; }
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q) if (q != c->rank && c->peers.p[q]) { cudaIpcCloseMemHandle(c->peers.p[q]); c->peers.p[q] = nullptr; }
            return;
        }
        c->peers.p[r] = (uint8_t *)p;
    }
    c->p2p = true;
}

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
    if (nranks > 1) {
        setup_p2p(c, ctx->stream);
        // the path must be the same on every rank: one more tiny gather of the outcome (a rank whose IPC mapping failed
        // would otherwise wait on NCCL while its peers wait on the mailboxes)
        int *d_f = nullptr;
        int mine = c->p2p ? 1 : 0;
        std::vector<int> all((size_t)nranks, 0);
        bool agreed = false;
        if (cudaMalloc((void **)&d_f, (size_t)(nranks + 1) * sizeof(int)) == cudaSuccess) {
            cudaMemcpyAsync(d_f + nranks, &mine, sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
            const int ge = g_api.AllGather(d_f + nranks, d_f, sizeof(int), 0, c->comm, ctx->stream);
            cudaMemcpyAsync(all.data(), d_f, (size_t)nranks * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
            agreed = ge == 0 && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
            cudaFree(d_f);
        }
        bool every = agreed;
        for (int r = 0; r < nranks && every; ++r) every = all[(size_t)r] == 1;
        if (!every) c->p2p = false;
        cudaGetLastError();
    }
    *out = c;
    return FXG_OK;
}

extern "C" int fxg_comm_uses_p2p(const fxg_comm *c) { return c && c->p2p ? 1 : 0; }

// after the stream has been synchronised: did a mailbox wait of this communicator time out (a peer never arrived)?
extern "C" int fxg_comm_check(fxg_comm *c) {
    if (c && c->p2p && c->h_status && *(volatile int *)c->h_status) {
        fxg_set_error("shard exchange timed out: a peer rank never published its block");
        return FXG_ECUDA;
    }
    return FXG_OK;
}

extern "C" int fxg_comm_nranks(const fxg_comm *c) { return c ? c->nranks : 1; }
extern "C" int fxg_comm_rank(const fxg_comm *c) { return c ? c->rank : 0; }

extern "C" void fxg_comm_destroy(fxg_comm *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->nranks && r < FXG_MBOX_MAXR; ++r)
        if (r != c->rank && c->peers.p[r]) cudaIpcCloseMemHandle(c->peers.p[r]);
    if (c->mbox) cudaFree(c->mbox);
    if (c->h_status) cudaFreeHost(c->h_status);
    if (c->comm && g_api.CommDestroy) g_api.CommDestroy(c->comm);
    cudaGetLastError();
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
    if (comm->p2p && bytes_per_rank <= FXG_MBOX_SLOT && bytes_per_rank % 16 == 0) {
        comm->epoch += 1;
        ctx->launches += 1;
        mbox_exchange_kernel<<<1, 256, 0, ctx->stream>>>(comm->peers, comm->nranks, comm->rank, comm->epoch, (const uint8_t *)d_send,
                                                         (uint8_t *)d_recv, (int)bytes_per_rank, comm->d_status);
        FXG_CUDA(cudaGetLastError());
        ctx->collectives += 1;
        return FXG_OK;
    }
    const int e = g_api.AllGather(d_send, d_recv, (size_t)bytes_per_rank, /*ncclInt8*/ 0, comm->comm, ctx->stream);
    if (e) return nccl_fail("ncclAllGather", e);
    ctx->collectives += 1;
    return FXG_OK;
}

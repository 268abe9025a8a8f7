// Multi-GPU exchange: ONE small ncclAllGather of the per-candidate partial tuples per
// candidate batch, then an identical rank-ordered merge on every rank.
// NCCL is dlopen'ed so that the library has no link-time dependency on it; single-GPU use
// never touches it.
#include <dlfcn.h>

#include <condition_variable>
#include <cstring>
#include <mutex>

#include "engine.hpp"

namespace mclcar {

int merge_tuples(int what, int F, int d, int G, int K, const double* in, double* out);

namespace {

// the few NCCL declarations we need (ABI-stable since NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat64 = 8;

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl(mclcar_ctx* ctx) {
    if (g_nccl.lib) return MCLCAR_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) {
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(ctx, MCLCAR_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    NcclApi api;
    api.lib = lib;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(lib, "ncclCommInitRank");
    api.AllGather = (decltype(api.AllGather))dlsym(lib, "ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(lib, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy || !api.GetErrorString)
        return fail(ctx, MCLCAR_ENCCL, "libnccl lacks a required symbol");
    g_nccl = api;
    return MCLCAR_OK;
}

}  // namespace

void comm_destroy(mclcar_ctx* ctx) {
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}

int nccl_all_gather_dev(mclcar_ctx* ctx, const double* d_send, double* d_recv, size_t cnt) {
    ncclResult_t r = g_nccl.AllGather(d_send, d_recv, cnt, kNcclFloat64, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != 0) return fail(ctx, MCLCAR_ENCCL, "ncclAllGather: %s", g_nccl.GetErrorString(r));
    return MCLCAR_OK;
}

// ---- in-process group: G contexts driven by G host threads of one process meet here instead of in NCCL ----
struct HostGroup {
    int G = 0, refs = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    uint64_t generation = 0;
    std::vector<double> buf[2];   // alternate buffers: a slow member may still merge round r while round r + 1 fills
    int failed = 0;
};

static int host_group_exchange(mclcar_ctx* ctx, int what, int F, int d, int K, double* tuples) {
    HostGroup* hg = (HostGroup*)ctx->host_group;
    const int G = hg->G;
    const int64_t T = tuple_len(what, F, d);
    const size_t cnt = (size_t)K * T;
    std::unique_lock<std::mutex> lk(hg->m);
    const uint64_t gen = hg->generation;
    std::vector<double>& b = hg->buf[gen & 1];
    if (hg->arrived == 0) b.assign((size_t)G * cnt, 0.0);
    if (b.size() != (size_t)G * cnt) hg->failed = 1;        // members disagree on the batch
    else std::memcpy(b.data() + (size_t)ctx->rank * cnt, tuples, cnt * sizeof(double));
    if (++hg->arrived == G) {
        hg->arrived = 0;
        ++hg->generation;
        hg->cv.notify_all();
    } else {
        hg->cv.wait(lk, [&] { return hg->generation != gen; });
    }
    const int failed = hg->failed;
    lk.unlock();
    if (failed) return fail(ctx, MCLCAR_ESTATE, "the contexts of the group were called with different candidate batches");
    return merge_tuples(what, F, d, G, K, b.data(), tuples);
}

void host_group_leave(mclcar_ctx* ctx) {
    HostGroup* hg = (HostGroup*)ctx->host_group;
    if (!hg) return;
    ctx->host_group = nullptr;
    bool last;
    {
        std::lock_guard<std::mutex> lk(hg->m);
        last = --hg->refs <= 0;
    }
    if (last) delete hg;
}

void* host_group_new(int G) {
    HostGroup* hg = new (std::nothrow) HostGroup();
    if (hg) hg->G = hg->refs = G;
    return hg;
}

// exchange of the K tuples with the other members of an in-process group, merged in rank order in place
int gather_and_merge(mclcar_ctx* ctx, int what, int F, int d, int K, double* tuples) {
    if (ctx->host_group) return host_group_exchange(ctx, what, F, d, K, tuples);
    // ranks of an NCCL job: run_reduce(merge_ranks = true) has gathered on the device and merged already
    return MCLCAR_OK;
}

}  // namespace mclcar

using namespace mclcar;

extern "C" {

int mclcar_nccl_unique_id(mclcar_ctx* ctx, unsigned char id_out[128]) {
    if (!ctx || !id_out) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    MCL_TRY(load_nccl(ctx));
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != 0) return fail(ctx, MCLCAR_ENCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    std::memcpy(id_out, id.internal, 128);
    return MCLCAR_OK;
}

int mclcar_nccl_comm_init(mclcar_ctx* ctx, const unsigned char id_in[128]) {
    if (!ctx || !id_in) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    MCL_TRY(load_nccl(ctx));
    if (ctx->world < 2) return fail(ctx, MCLCAR_ESTATE, "call mclcar_ctx_set_shard(rank, world >= 2) first");
    comm_destroy(ctx);
    ncclUniqueId id;
    std::memcpy(id.internal, id_in, 128);
    ncclComm_t comm = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&comm, ctx->world, id, ctx->rank);
    if (r != 0) return fail(ctx, MCLCAR_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    ctx->nccl_comm = comm;
    return MCLCAR_OK;
}

}  // extern "C"

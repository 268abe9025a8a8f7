// Multi-GPU exchange: ONE small ncclAllGather of the per-candidate partial tuples per
// candidate batch, then an identical rank-ordered merge on every rank.
// NCCL is dlopen'ed so that the library has no link-time dependency on it; single-GPU use
// never touches it.
#include <dlfcn.h>

#include <cstring>

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

// all-gather the K tuples of every rank and merge them in rank order; in place on `tuples`
int gather_and_merge(mclcar_ctx* ctx, int what, int F, int d, int K, double* tuples) {
    if (!ctx->nccl_comm || ctx->world <= 1) return MCLCAR_OK;
    const int G = ctx->world;
    const int64_t T = tuple_len(what, F, d);
    const size_t cnt = (size_t)K * T;
    MCL_TRY(ensure(ctx, ctx->d_gather, (size_t)(G + 1) * cnt * sizeof(double)));
    double* send = (double*)ctx->d_gather.p;
    double* recv = send + cnt;
    cudaStream_t st = ctx->stream;
    MCL_CUDA(ctx, cudaMemcpyAsync(send, tuples, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
    ncclResult_t r = g_nccl.AllGather(send, recv, cnt, kNcclFloat64, (ncclComm_t)ctx->nccl_comm, st);
    if (r != 0) return fail(ctx, MCLCAR_ENCCL, "ncclAllGather: %s", g_nccl.GetErrorString(r));
    std::vector<double> all((size_t)G * cnt);
    MCL_CUDA(ctx, cudaMemcpyAsync(all.data(), recv, all.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    MCL_CUDA(ctx, cudaStreamSynchronize(st));
    return merge_tuples(what, F, d, G, K, all.data(), tuples);
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

// Context, model upload (graph / design / observations) and small host utilities.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>

#include "engine.hpp"

namespace mclcar {

static thread_local std::string g_last_error;

void set_global_error(const char* msg) { g_last_error = msg; }

int fail(mclcar_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_last_error = buf;
    return code;
}

int require_device(mclcar_ctx* ctx) {
    if (!ctx) return MCLCAR_EINVAL;
    if (ctx->host_only) return fail(ctx, MCLCAR_ENODEV, "compute call on a host-only context: there is no CPU fallback");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, MCLCAR_ECUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
    // a non-sticky error left behind by an earlier, unrelated runtime call (ours or another library's in
    // this process) must not be blamed on this call's first launch check
    cudaGetLastError();
    return MCLCAR_OK;
}

// best-fit block of the cache (at most 25 % larger than asked), else a fresh cudaMalloc; when
// that fails every cached free block is returned to the driver and the allocation retried
int pool_alloc(mclcar_ctx* ctx, void** out, size_t bytes) {
    *out = nullptr;
    if (bytes == 0) bytes = 16;
    int best = -1;
    for (int i = 0; i < (int)ctx->pool.size(); ++i) {
        auto& b = ctx->pool[i];
        if (!b.used && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 + 4096 &&
            (best < 0 || b.bytes < ctx->pool[best].bytes))
            best = i;
    }
    if (best >= 0) {
        ctx->pool[best].used = true;
        *out = ctx->pool[best].p;
        return MCLCAR_OK;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaStreamSynchronize(ctx->stream);
        for (auto it = ctx->pool.begin(); it != ctx->pool.end();) {
            if (!it->used) { cudaFree(it->p); it = ctx->pool.erase(it); } else ++it;
        }
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, MCLCAR_ENOMEM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    ctx->pool.push_back({p, bytes, true});
    *out = p;
    return MCLCAR_OK;
}

// the block becomes reusable by later calls on the same stream (stream order makes that safe)
void pool_free(mclcar_ctx* ctx, void* p) {
    if (!p) return;
    for (auto& b : ctx->pool)
        if (b.p == p) {
            b.used = false;
            // never sit on more than half of the device: a session that changes problem sizes would
            // otherwise accumulate blocks of every size it has ever used
            pool_trim(ctx, ctx->pool_cap);
            return;
        }
    cudaFree(p);  // not ours
}

// return cached FREE blocks to the driver, largest first, until at most keep_bytes stay cached
void pool_trim(mclcar_ctx* ctx, size_t keep_bytes) {
    size_t cached = 0;
    for (auto& b : ctx->pool)
        if (!b.used) cached += b.bytes;
    if (cached <= keep_bytes) return;
    cudaStreamSynchronize(ctx->stream);
    while (cached > keep_bytes) {
        int big = -1;
        for (int i = 0; i < (int)ctx->pool.size(); ++i)
            if (!ctx->pool[i].used && (big < 0 || ctx->pool[i].bytes > ctx->pool[big].bytes)) big = i;
        if (big < 0) break;
        cached -= ctx->pool[big].bytes;
        cudaFree(ctx->pool[big].p);
        ctx->pool.erase(ctx->pool.begin() + big);
    }
}

void pool_release_all(mclcar_ctx* ctx) {
    for (auto& b : ctx->pool) cudaFree(b.p);
    ctx->pool.clear();
}

int ensure(mclcar_ctx* ctx, DeviceBuf& b, size_t bytes) {
    if (bytes <= b.bytes && b.p) return MCLCAR_OK;
    if (b.p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(b.p);
        b.p = nullptr;
        b.bytes = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    MCL_CUDA(ctx, cudaMalloc(&b.p, want));
    b.bytes = want;
    return MCLCAR_OK;
}

int ensure_pinned(mclcar_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->h_pinned_bytes && ctx->h_pinned) return MCLCAR_OK;
    if (ctx->h_pinned) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeHost(ctx->h_pinned);
        ctx->h_pinned = nullptr;
        ctx->h_pinned_bytes = 0;
    }
    size_t want = bytes + bytes / 4 + 4096;
    MCL_CUDA(ctx, cudaMallocHost(&ctx->h_pinned, want));
    ctx->h_pinned_bytes = want;
    return MCLCAR_OK;
}

// MCLCAR_TRACE_HOST=1: wall-clock sections of the host-side set-up calls on stderr (scripts/exp_host_algebra.py)
struct HostTrace {
    const char* what;
    bool on;
    std::chrono::steady_clock::time_point t;
    explicit HostTrace(const char* w) : what(w), on(getenv("MCLCAR_TRACE_HOST") != nullptr), t(std::chrono::steady_clock::now()) {}
    void mark(const char* section) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[host] %s: %s %.3f ms\n", what, section, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

ProfScope::ProfScope(mclcar_ctx* c, int category, int64_t n_launches) : ctx(c), cat(category) {
    ctx->launches[cat] += n_launches;
    if (ctx->profiling && cudaEventCreate(&a) == cudaSuccess) cudaEventRecord(a, ctx->stream);
}
void ProfScope::close() {
    if (!a) return;
    cudaEvent_t b = nullptr;
    if (cudaEventCreate(&b) == cudaSuccess) {
        cudaEventRecord(b, ctx->stream);
        ctx->spans.push_back({cat, a, b});
    } else {
        cudaEventDestroy(a);
    }
    a = nullptr;
}

template <class T>
static int upload(mclcar_ctx* ctx, T*& dptr, const std::vector<T>& h) {
    if (ctx->host_only) return MCLCAR_OK;
    if (dptr) {
        pool_free(ctx, dptr);
        dptr = nullptr;
    }
    if (h.empty()) return MCLCAR_OK;
    MCL_TRY(pool_alloc(ctx, (void**)&dptr, h.size() * sizeof(T)));
    MCL_CUDA(ctx, cudaMemcpyAsync(dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCLCAR_OK;
}

template <typename T>
static int upload_from(mclcar_ctx* ctx, T*& dptr, const T* src, size_t count) {   // straight from the caller's buffer (fast when it is pinned)
    if (ctx->host_only) return MCLCAR_OK;
    if (dptr) {
        pool_free(ctx, dptr);
        dptr = nullptr;
    }
    if (!count) return MCLCAR_OK;
    MCL_TRY(pool_alloc(ctx, (void**)&dptr, count * sizeof(T)));
    MCL_CUDA(ctx, cudaMemcpyAsync(dptr, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCLCAR_OK;
}

// The host algebra of set_design / set_obs (W X, X'X, X'WX, the statistics of y) is O(n) per call and sits in every
// end-to-end step: large lattices split it over a FIXED number of index ranges (results do not depend on the number
// of host threads that happen to run them), each range on its own thread.
constexpr int kHostParts = 8;
template <class Fn>
static void host_parts(int n, Fn fn, int64_t work = -1) {   // fn(part, begin, end); work: elements touched (default n)
    if ((work < 0 ? (int64_t)n : work) < (1 << 17)) {
        for (int k = 0; k < kHostParts; ++k) fn(k, (int)((int64_t)n * k / kHostParts), (int)((int64_t)n * (k + 1) / kHostParts));
        return;
    }
    std::vector<std::thread> th;
    for (int k = 0; k < kHostParts; ++k)
        th.emplace_back([=] { fn(k, (int)((int64_t)n * k / kHostParts), (int)((int64_t)n * (k + 1) / kHostParts)); });
    for (auto& t : th) t.join();
}

// dot product: four independent accumulators per index range (vectorises; error ~ sqrt(n) eps), ranges added in order
static double dot4(const double* a, const double* b, int n) {
    double part[kHostParts];
    host_parts(n, [&](int k, int i0, int i1) {
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int i = i0;
        for (; i + 3 < i1; i += 4) {
            s0 += a[i] * b[i];
            s1 += a[i + 1] * b[i + 1];
            s2 += a[i + 2] * b[i + 2];
            s3 += a[i + 3] * b[i + 3];
        }
        for (; i < i1; ++i) s0 += a[i] * b[i];
        part[k] = (s0 + s1) + (s2 + s3);
    });
    double s = 0;
    for (int k = 0; k < kHostParts; ++k) s += part[k];
    return s;
}

// y = W x on the host: the rook torus as a stencil (no index loads), anything else through the CSR copy
static void spmv(const mclcar_ctx* c, const double* x, double* y) {
    if (c->kind == GRAPH_TORUS) {
        const int n1 = c->n1, n2 = c->n2;
        host_parts(n1, [&](int, int r0, int r1) {
            for (int i = r0; i < r1; ++i) {
                const double* row = x + (size_t)i * n2;
                const double* up = x + (size_t)(i == 0 ? n1 - 1 : i - 1) * n2;
                const double* dn = x + (size_t)(i + 1 == n1 ? 0 : i + 1) * n2;
                double* o = y + (size_t)i * n2;
                // same association as the CSR row (columns in increasing order) is not needed: W x is exact to rounding
                // either way, and every consumer takes it through tolerance-checked sums
                for (int j = 0; j < n2; ++j) {
                    const int jl = j == 0 ? n2 - 1 : j - 1, jr = j + 1 == n2 ? 0 : j + 1;
                    o[j] = (row[jl] + row[jr]) + (up[j] + dn[j]);
                }
            }
        }, (int64_t)n1 * n2);
        return;
    }
    host_parts(c->n, [&](int, int i0, int i1) {
        for (int i = i0; i < i1; ++i) {
            double a = 0.0;
            for (int e = c->rowptr[i]; e < c->rowptr[i + 1]; ++e) a += (c->binary ? 1.0 : c->val[e]) * x[c->colidx[e]];
            y[i] = a;
        }
    });
}

static void invalidate_model(mclcar_ctx* ctx) {
    ctx->has_design = false;
    ctx->has_obs = false;
    ctx->p = 0;
    ctx->X.clear(); ctx->WX.clear(); ctx->G0.clear(); ctx->G1.clear();
    ctx->y.clear(); ctx->ntrial.clear(); ctx->qobs.clear();
    ctx->design_pending = ctx->obs_pending = false;
}

// The O(n) host algebra behind set_design / set_obs, formed on first use.  W X goes to the device through a pinned
// staging buffer on the context's stream without a host wait (the device copy is consumed in stream order; the
// staging buffer is reused only after the copy's event).
int ensure_algebra(mclcar_ctx* ctx) {
    if (!ctx->design_pending && !ctx->obs_pending) return MCLCAR_OK;
    const int n = ctx->n, p = ctx->p;
    if (ctx->design_pending) {
        HostTrace tr("design algebra");
        ctx->WX.resize((size_t)n * p);
        for (int a = 0; a < p; ++a) spmv(ctx, ctx->X.data() + (size_t)a * n, ctx->WX.data() + (size_t)a * n);
        tr.mark("W X");
        ctx->G0.assign((size_t)p * p, 0.0);
        ctx->G1.assign((size_t)p * p, 0.0);
        for (int a = 0; a < p; ++a)
            for (int b = 0; b < p; ++b) {
                const double *xa = ctx->X.data() + (size_t)a * n, *xb = ctx->X.data() + (size_t)b * n,
                             *wb = ctx->WX.data() + (size_t)b * n;
                ctx->G0[(size_t)a * p + b] = dot4(xa, xb, n);
                ctx->G1[(size_t)a * p + b] = dot4(xa, wb, n);
            }
        tr.mark("X'X, X'WX");
        if (!ctx->host_only) {
            if (ctx->d_WX) {
                pool_free(ctx, ctx->d_WX);
                ctx->d_WX = nullptr;
            }
            const size_t bytes = (size_t)n * p * sizeof(double);
            if (bytes) {
                MCL_TRY(pool_alloc(ctx, (void**)&ctx->d_WX, bytes));
                if (ctx->wx_event) MCL_CUDA(ctx, cudaEventSynchronize(ctx->wx_event));   // the previous copy has left the staging buffer
                if (ctx->h_wx_bytes < bytes) {
                    if (ctx->h_wx) cudaFreeHost(ctx->h_wx);
                    ctx->h_wx = nullptr; ctx->h_wx_bytes = 0;
                    MCL_CUDA(ctx, cudaMallocHost(&ctx->h_wx, bytes));
                    ctx->h_wx_bytes = bytes;
                }
                double* stage = (double*)ctx->h_wx;
                const double* src = ctx->WX.data();
                host_parts((int)std::min<size_t>((size_t)n * p, (size_t)INT32_MAX), [&](int, int i0, int i1) {
                    std::memcpy(stage + i0, src + i0, (size_t)(i1 - i0) * sizeof(double));
                });
                MCL_CUDA(ctx, cudaMemcpyAsync(ctx->d_WX, stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
                if (!ctx->wx_event) MCL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->wx_event, cudaEventDisableTiming));
                MCL_CUDA(ctx, cudaEventRecord(ctx->wx_event, ctx->stream));
            }
            tr.mark("W X to the device (queued)");
        }
        ctx->design_pending = false;
    }
    if (ctx->obs_pending) {
        HostTrace tr("observation algebra");
        std::vector<double>& Wy = ctx->h_scratch;   // kept across calls: no 8 MB allocation and page faults per step
        Wy.resize(n);
        const double* y = ctx->y.data();
        spmv(ctx, y, Wy.data());
        tr.mark("W y");
        ctx->qobs.assign(2 + 2 * p, 0.0);
        ctx->qobs[0] = dot4(y, y, n);
        ctx->qobs[1] = dot4(y, Wy.data(), n);
        for (int a = 0; a < p; ++a) {
            const double* xa = ctx->X.data() + (size_t)a * n;
            ctx->qobs[2 + a] = dot4(xa, y, n);
            ctx->qobs[2 + p + a] = dot4(xa, Wy.data(), n);
        }
        tr.mark("y'y, y'Wy, X'y, X'Wy");
        ctx->obs_pending = false;
    }
    return MCLCAR_OK;
}

// greedy colouring in natural order (smallest free colour); colour classes keep site order
static void colour_graph(mclcar_ctx* c) {
    const int n = c->n;
    c->color_of.assign(n, -1);
    int nc = 0;
    if (c->kind == GRAPH_TORUS && c->n1 % 2 == 0 && c->n2 % 2 == 0) {
        for (int i = 0; i < c->n1; ++i)
            for (int j = 0; j < c->n2; ++j) c->color_of[i * c->n2 + j] = (i + j) & 1;
        nc = 2;
    } else {
        std::vector<int> mark(c->max_deg + 2, -1);
        for (int i = 0; i < n; ++i) {
            for (int e = c->rowptr[i]; e < c->rowptr[i + 1]; ++e) {
                const int cj = c->color_of[c->colidx[e]];
                if (cj >= 0 && cj < (int)mark.size()) mark[cj] = i;
            }
            int col = 0;
            while (mark[col] == i) ++col;
            c->color_of[i] = col;
            nc = std::max(nc, col + 1);
        }
    }
    c->ncolors = nc;
    c->color_ptr.assign(nc + 1, 0);
    for (int i = 0; i < n; ++i) c->color_ptr[c->color_of[i] + 1]++;
    for (int k = 0; k < nc; ++k) c->color_ptr[k + 1] += c->color_ptr[k];
    c->order.resize(n);
    std::vector<int> pos(c->color_ptr.begin(), c->color_ptr.end() - 1);
    for (int i = 0; i < n; ++i) c->order[pos[c->color_of[i]]++] = i;
}

void edge_cache_clear(mclcar_ctx* ctx) {
    for (auto& c : ctx->edge_cache) {
        void* dev[] = {c.d_eptr, c.d_ei, c.d_ej, c.d_ew, c.d_diag};
        for (void* d : dev)
            if (d && !ctx->host_only) pool_free(ctx, d);
    }
    ctx->edge_cache.clear();
}

static int finish_graph(mclcar_ctx* ctx) {
    edge_cache_clear(ctx);   // cached edge lists are keyed by the host CSR arrays that just changed
    ctx->max_deg = 0;
    for (int i = 0; i < ctx->n; ++i) ctx->max_deg = std::max(ctx->max_deg, ctx->rowptr[i + 1] - ctx->rowptr[i]);
    colour_graph(ctx);
    MCL_TRY(upload(ctx, ctx->d_rowptr, ctx->rowptr));
    MCL_TRY(upload(ctx, ctx->d_colidx, ctx->colidx));
    MCL_TRY(upload(ctx, ctx->d_val, ctx->val));
    MCL_TRY(upload(ctx, ctx->d_order, ctx->order));
    return MCLCAR_OK;
}

}  // namespace mclcar

using namespace mclcar;

extern "C" {

int mclcar_ctx_create(int device, uint64_t seed, mclcar_ctx** out) {
    if (!out) return MCLCAR_EINVAL;
    *out = nullptr;
    mclcar_ctx* ctx = new (std::nothrow) mclcar_ctx();
    if (!ctx) return MCLCAR_ENOMEM;
    ctx->seed = seed;
    ctx->device = device;
    ctx->host_only = device < 0;
    if (!ctx->host_only) {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || device >= ndev) {
            fail(nullptr, MCLCAR_ENODEV, "CUDA device %d not available (%s, %d devices): no CPU fallback", device,
                 e == cudaSuccess ? "ok" : cudaGetErrorString(e), ndev);
            delete ctx;
            return MCLCAR_ENODEV;
        }
        cudaDeviceProp prop;
        if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
            fail(nullptr, MCLCAR_ECUDA, "device %d initialisation failed: %s", device, cudaGetErrorString(e));
            delete ctx;
            return MCLCAR_ECUDA;
        }
        if (prop.major != 10) {
            fail(nullptr, MCLCAR_ENODEV, "device %d is sm_%d%d; this library contains sm_100a code only", device,
                 prop.major, prop.minor);
            cudaStreamDestroy(ctx->own_stream);
            delete ctx;
            return MCLCAR_ENODEV;
        }
        ctx->n_sm = prop.multiProcessorCount;
        ctx->pool_cap = prop.totalGlobalMem / 2;
        ctx->stream = ctx->own_stream;
    }
    *out = ctx;
    return MCLCAR_OK;
}

void mclcar_ctx_destroy(mclcar_ctx* ctx) {
    if (!ctx) return;
    hcar_forget(ctx);
    if (!ctx->host_only) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        void* ptrs[] = {ctx->d_coef.p, ctx->d_partial.p, ctx->d_tuples.p, ctx->d_subidx.p,
                        ctx->d_gather.p, ctx->d_sptab.p};
        for (void* p : ptrs)
            if (p) cudaFree(p);
        pool_release_all(ctx);
        comm_destroy(ctx);
        for (auto& sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
        if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
        if (ctx->h_wx) cudaFreeHost(ctx->h_wx);
        if (ctx->wx_event) cudaEventDestroy(ctx->wx_event);
        if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    }
    delete ctx;
}

const char* mclcar_last_error(const mclcar_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int mclcar_ctx_set_stream(mclcar_ctx* ctx, void* cuda_stream) {
    if (!ctx) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return MCLCAR_OK;
}

int mclcar_ctx_trim(mclcar_ctx* ctx) {
    if (!ctx) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    pool_trim(ctx, 0);
    return MCLCAR_OK;
}

int mclcar_ctx_sync(mclcar_ctx* ctx) {
    if (!ctx) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCLCAR_OK;
}

int mclcar_ctx_set_shard(mclcar_ctx* ctx, int rank, int world) {
    if (!ctx) return MCLCAR_EINVAL;
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, MCLCAR_EINVAL, "bad shard %d of %d", rank, world);
    ctx->rank = rank;
    ctx->world = world;
    return MCLCAR_OK;
}

int mclcar_ctx_set_draw(mclcar_ctx* ctx, uint64_t draw) {
    if (!ctx) return MCLCAR_EINVAL;
    ctx->draw = draw;
    return MCLCAR_OK;
}

int mclcar_ctx_get_draw(const mclcar_ctx* ctx, uint64_t* draw) {
    if (!ctx || !draw) return MCLCAR_EINVAL;
    *draw = ctx->draw;
    return MCLCAR_OK;
}

int mclcar_ctx_profile(mclcar_ctx* ctx, int enable) {
    if (!ctx) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    cudaStreamSynchronize(ctx->stream);
    for (auto& sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    ctx->spans.clear();
    for (int c = 0; c < PROF_NCAT; ++c) { ctx->launches[c] = 0; ctx->span_ms[c] = 0.0; }
    ctx->profiling = enable != 0;
    return MCLCAR_OK;
}

int mclcar_ctx_profile_read(mclcar_ctx* ctx, double ms_out[8], int64_t launches_out[8]) {
    if (!ctx) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& sp : ctx->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) ctx->span_ms[sp.cat] += ms;
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    ctx->spans.clear();
    for (int c = 0; c < PROF_NCAT; ++c) {
        if (ms_out) ms_out[c] = ctx->span_ms[c];
        if (launches_out) launches_out[c] = ctx->launches[c];
    }
    return MCLCAR_OK;
}

}  // extern "C"

namespace mclcar {
// The reference recomputes eigen(W) / chol.spam inside every call when the caller gave no spectrum (R/mcl.bin.R:176,
// R/mcl.HCAR.R:166-173).  Here: estimate the spectral measure once per graph on the device (spectrum.cu).
int ensure_spectrum(mclcar_ctx* ctx) {
    if (!ctx->eigs.empty()) return MCLCAR_OK;
    if (ctx->kind == GRAPH_NONE) return fail(ctx, MCLCAR_ESTATE, "set the graph first");
    if (ctx->host_only)
        return fail(ctx, MCLCAR_ESTATE, "the eigenvalues of W are needed (mclcar_graph_set_eigenvalues): a host-only context cannot estimate them");
    if (ctx->rowptr[ctx->n] == 0) return MCLCAR_OK;    // empty graph (hierarchical model without W): nothing to sum
    return mclcar_graph_estimate_spectrum(ctx, 0, ctx->seed);
}
}  // namespace mclcar

extern "C" {

int mclcar_graph_n(const mclcar_ctx* ctx) { return ctx ? ctx->n : 0; }
int mclcar_graph_ncolors(const mclcar_ctx* ctx) { return ctx ? ctx->ncolors : 0; }

static int set_torus(mclcar_ctx* ctx, int n1, int n2) {
    invalidate_model(ctx);
    ctx->kind = GRAPH_TORUS;
    ctx->n1 = n1; ctx->n2 = n2; ctx->n = n1 * n2;
    ctx->binary = true;
    ctx->val.clear();
    const int n = ctx->n;
    ctx->rowptr.resize(n + 1);
    ctx->colidx.resize((size_t)4 * n);
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j) {
            const int s = i * n2 + j;
            int nb[4] = {((i + n1 - 1) % n1) * n2 + j, i * n2 + (j + n2 - 1) % n2, i * n2 + (j + 1) % n2,
                         ((i + 1) % n1) * n2 + j};
            std::sort(nb, nb + 4);
            ctx->rowptr[s] = 4 * s;
            for (int t = 0; t < 4; ++t) ctx->colidx[(size_t)4 * s + t] = nb[t];
        }
    ctx->rowptr[n] = 4 * n;
    // analytic spectrum: 2cos(2 pi j/n1) + 2cos(2 pi k/n2) (replaces eigen(W), R/CAR.simLM.R:8)
    ctx->eig_w.clear();
    ctx->eigs.resize(n);
    std::vector<double> a(n1), b(n2);
    for (int j = 0; j < n1; ++j) a[j] = 2.0 * std::cos(2.0 * M_PI * j / n1);
    for (int k = 0; k < n2; ++k) b[k] = 2.0 * std::cos(2.0 * M_PI * k / n2);
    for (int j = 0; j < n1; ++j)
        for (int k = 0; k < n2; ++k) ctx->eigs[(size_t)j * n2 + k] = a[j] + b[k];
    ctx->ew_min = *std::min_element(ctx->eigs.begin(), ctx->eigs.end());
    ctx->ew_max = *std::max_element(ctx->eigs.begin(), ctx->eigs.end());
    return finish_graph(ctx);
}

int mclcar_graph_torus(mclcar_ctx* ctx, int n1, int n2) {
    if (!ctx) return MCLCAR_EINVAL;
    if (n1 < 3 || n2 < 3 || (int64_t)n1 * n2 > (int64_t)1 << 30)
        return fail(ctx, MCLCAR_EINVAL, "torus sides must be >= 3 and n1*n2 <= 2^30 (got %d x %d)", n1, n2);
    if (!ctx->host_only) MCL_TRY(require_device(ctx));
    return set_torus(ctx, n1, n2);
}

// Is this binary CSR matrix exactly the rook torus with site (i, j) at index i*n2 + j?  Every row has four
// distinct neighbours; site 0's are {1, n2-1, n2, n-n2}, so n2 is its third smallest neighbour (n1, n2 >= 3;
// sides of length 4 or less give coinciding candidates that the full comparison below sorts out).
static bool detect_rook_torus(int n, const int* rowptr, const int* colidx, const double* val, int* n1_out, int* n2_out) {
    if (n < 9 || rowptr[n] != (int64_t)4 * n) return false;
    for (int i = 0; i < n; ++i)
        if (rowptr[i + 1] - rowptr[i] != 4) return false;
    if (val)
        for (int64_t e = 0; e < (int64_t)4 * n; ++e)
            if (val[e] != 1.0) return false;
    int nb0[4] = {colidx[0], colidx[1], colidx[2], colidx[3]};
    std::sort(nb0, nb0 + 4);
    // candidates for n2: any neighbour index of site 0 (+1) that divides n -- at most a handful
    for (int cand : {nb0[2], nb0[1] + 1, nb0[1], nb0[0] + 1, nb0[3]}) {
        const int n2 = cand;
        if (n2 < 3 || n % n2 != 0 || n / n2 < 3) continue;
        const int n1 = n / n2;
        bool ok = true;
        for (int i = 0; i < n1 && ok; ++i)
            for (int j = 0; j < n2 && ok; ++j) {
                const int s = i * n2 + j;
                int want[4] = {((i + n1 - 1) % n1) * n2 + j, i * n2 + (j + n2 - 1) % n2, i * n2 + (j + 1) % n2,
                               ((i + 1) % n1) * n2 + j};
                int got[4] = {colidx[4 * (size_t)s], colidx[4 * (size_t)s + 1], colidx[4 * (size_t)s + 2], colidx[4 * (size_t)s + 3]};
                std::sort(want, want + 4);
                std::sort(got, got + 4);
                ok = want[0] == got[0] && want[1] == got[1] && want[2] == got[2] && want[3] == got[3] &&
                     want[0] != want[1] && want[1] != want[2] && want[2] != want[3];
            }
        if (ok) { *n1_out = n1; *n2_out = n2; return true; }
    }
    return false;
}

int mclcar_graph_csr_opts(mclcar_ctx* ctx, int n, const int* rowptr, const int* colidx, const double* val, int detect_torus) {
    if (!ctx || !rowptr || !colidx) return MCLCAR_EINVAL;
    if (n < 1) return fail(ctx, MCLCAR_EINVAL, "n must be positive");
    if (!ctx->host_only) MCL_TRY(require_device(ctx));
    if (rowptr[0] != 0) return fail(ctx, MCLCAR_EINVAL, "rowptr[0] must be 0");
    const int64_t nnz = rowptr[n];
    for (int i = 0; i < n; ++i)
        if (rowptr[i + 1] < rowptr[i]) return fail(ctx, MCLCAR_EINVAL, "rowptr not monotone at row %d", i);
    // validate: indices in range, zero diagonal, symmetric (W is used as a symmetric matrix by
    // the reference without checking: eigen(W, symmetric = TRUE), R/loglik.dCAR.R:25)
    for (int i = 0; i < n; ++i)
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const int j = colidx[e];
            if (j < 0 || j >= n) return fail(ctx, MCLCAR_EINVAL, "column index %d out of range in row %d", j, i);
            if (j == i) return fail(ctx, MCLCAR_EINVAL, "W must have a zero diagonal (row %d)", i);
            for (int e2 = rowptr[i]; e2 < e; ++e2)
                if (colidx[e2] == j) return fail(ctx, MCLCAR_EINVAL, "duplicate entry (%d, %d)", i, j);
        }
    for (int i = 0; i < n; ++i)
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {   // the mirror entry (j, i) must exist with the same value
            const int j = colidx[e];
            bool found = false;
            for (int e2 = rowptr[j]; e2 < rowptr[j + 1] && !found; ++e2)
                found = colidx[e2] == i && (val ? val[e2] == val[e] : true);
            if (!found) return fail(ctx, MCLCAR_EINVAL, "W is not symmetric at (%d, %d)", i, j);
        }
    int t1 = 0, t2 = 0;
    if (detect_torus && detect_rook_torus(n, rowptr, colidx, val, &t1, &t2)) return set_torus(ctx, t1, t2);
    invalidate_model(ctx);
    ctx->kind = GRAPH_CSR;
    ctx->n = n; ctx->n1 = ctx->n2 = 0;
    ctx->rowptr.assign(rowptr, rowptr + n + 1);
    ctx->colidx.assign(colidx, colidx + nnz);
    ctx->binary = true;
    if (val)
        for (int64_t e = 0; e < nnz; ++e)
            if (val[e] != 1.0) ctx->binary = false;
    if (ctx->binary) ctx->val.clear();
    else ctx->val.assign(val, val + nnz);
    ctx->eigs.clear();
    ctx->eig_w.clear();
    ctx->ew_min = ctx->ew_max = 0.0;
    return finish_graph(ctx);
}

int mclcar_graph_csr(mclcar_ctx* ctx, int n, const int* rowptr, const int* colidx, const double* val) {
    return mclcar_graph_csr_opts(ctx, n, rowptr, colidx, val, 1);
}

int mclcar_graph_kind(const mclcar_ctx* ctx, int* n1, int* n2) {
    if (!ctx) return 0;
    if (n1) *n1 = ctx->n1;
    if (n2) *n2 = ctx->n2;
    return ctx->kind;
}

int mclcar_graph_set_eigenvalues(mclcar_ctx* ctx, const double* ew, int n) {
    if (!ctx || !ew) return MCLCAR_EINVAL;
    if (ctx->kind == GRAPH_NONE) return fail(ctx, MCLCAR_ESTATE, "set the graph first");
    if (n != ctx->n) return fail(ctx, MCLCAR_EINVAL, "expected %d eigenvalues, got %d", ctx->n, n);
    ctx->eigs.assign(ew, ew + n);
    ctx->eig_w.clear();
    ctx->ew_min = *std::min_element(ctx->eigs.begin(), ctx->eigs.end());
    ctx->ew_max = *std::max_element(ctx->eigs.begin(), ctx->eigs.end());
    return MCLCAR_OK;
}

int mclcar_set_design(mclcar_ctx* ctx, const double* X, int p) {
    if (!ctx) return MCLCAR_EINVAL;
    if (ctx->kind == GRAPH_NONE) return fail(ctx, MCLCAR_ESTATE, "set the graph before the design");
    if (p < 0 || p > MCLCAR_MAX_COV) return fail(ctx, MCLCAR_ELIMIT, "p = %d outside [0, %d]", p, MCLCAR_MAX_COV);
    if (p > 0 && !X) return MCLCAR_EINVAL;
    if (!ctx->host_only) MCL_TRY(require_device(ctx));
    const int n = ctx->n;
    HostTrace tr("set_design");
    ctx->p = p;
    ctx->X.resize((size_t)n * p);
    // the host copy, and in the same pass: is every column one constant?  (then X beta is one number and the
    // statistics of the design columns follow from sum x)
    bool all_const = true;
    for (int a = 0; a < p; ++a) {
        const double* src = X + (size_t)a * n;
        double* dst = ctx->X.data() + (size_t)a * n;
        bool part_const[kHostParts];
        host_parts(n, [&](int k, int i0, int i1) {
            const double x0 = src[0];
            bool c = true;
            for (int i = i0; i < i1; ++i) {
                dst[i] = src[i];
                c &= src[i] == x0;
            }
            part_const[k] = c;
        });
        for (int k = 0; k < kHostParts; ++k) all_const = all_const && part_const[k];
    }
    ctx->design_const = all_const;
    tr.mark("copy X, constant columns");
    MCL_TRY(upload_from(ctx, ctx->d_X, X, (size_t)n * p));
    tr.mark("upload X");
    // W X, X'X, X'WX: formed when first needed (ensure_algebra) -- in an end-to-end step that is after the sampler's
    // launches have been queued, so the host works while the device sweeps
    ctx->design_pending = true;
    ctx->has_design = true;
    ctx->has_obs = false;  // qobs depends on X
    ctx->obs_pending = false;
    return MCLCAR_OK;
}

int mclcar_set_obs(mclcar_ctx* ctx, const double* y, const double* ntrial) {
    if (!ctx || !y) return MCLCAR_EINVAL;
    if (ctx->kind == GRAPH_NONE) return fail(ctx, MCLCAR_ESTATE, "set the graph before the observations");
    if (!ctx->has_design) {  // no regression part: p = 0
        MCL_TRY(mclcar_set_design(ctx, nullptr, 0));
    }
    if (!ctx->host_only) MCL_TRY(require_device(ctx));
    const int n = ctx->n;
    HostTrace tr("set_obs");
    ctx->y.resize(n);
    {
        double* dst = ctx->y.data();
        host_parts(n, [&](int, int i0, int i1) { std::memcpy(dst + i0, y + i0, (size_t)(i1 - i0) * sizeof(double)); });
    }
    if (ntrial) ctx->ntrial.assign(ntrial, ntrial + n);
    else ctx->ntrial.clear();
    tr.mark("copy y");
    MCL_TRY(upload_from(ctx, ctx->d_y, y, (size_t)n));
    MCL_TRY(upload(ctx, ctx->d_ntrial, ctx->ntrial));
    tr.mark("upload y");
    ctx->obs_pending = true;   // y'y, y'Wy, X'y, X'Wy: ensure_algebra
    ctx->has_obs = true;
    return MCLCAR_OK;
}

}  // extern "C"

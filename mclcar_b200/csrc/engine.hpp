// Internal declarations shared by the translation units of libmclcar_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mclcar_b200.h"

namespace mclcar {

enum GraphKind { GRAPH_NONE = 0, GRAPH_TORUS = 1, GRAPH_CSR = 2 };

// tuple kinds ("what" at the ABI)
enum { WHAT_LR = 0, WHAT_GRAD = 1, WHAT_HESS = 2 };

// instrumentation categories (mclcar_ctx_profile_read)
enum { PROF_SWEEP = 0, PROF_EMIT = 1, PROF_STATS = 2, PROF_REDUCE = 3, PROF_CSR_SAMPLER = 4, PROF_GLM = 5, PROF_SPECTRUM = 6, PROF_NCAT = 8 };

// upper-triangle edge list of a symmetric CSR matrix, on the device (stats_csr.cu), cached per matrix
struct EdgeCache {
    const int* key_rowptr = nullptr;
    const int* key_colidx = nullptr;
    int n = 0, nnz = 0;
    int *d_eptr = nullptr, *d_ei = nullptr, *d_ej = nullptr;
    double *d_ew = nullptr, *d_diag = nullptr;
};

struct DeviceBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

}  // namespace mclcar

struct mclcar_ctx {
    int device = -1;
    bool host_only = true;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    uint64_t seed = 0;
    // every *_prep call owns a fresh Philox stream: key = seed + draw * golden ratio (begin_draw); the
    // counter is part of the reproducible state (same seed + same call index => same samples) and is
    // the same on every rank of a sharded run
    uint64_t draw = 0;       // index of the next prep call
    uint64_t draw_salt = 0;  // > 0: a repeated run inside one prep call (pilot-checked GLM thinning)
    uint64_t draw_key = 0;   // Philox key of the prep call in progress
    int rank = 0, world = 1;
    int n_sm = 148;
    std::string err;

    // ---- graph (host copy is always CSR; the torus keeps n1, n2 as well)
    int kind = mclcar::GRAPH_NONE;
    int n = 0, n1 = 0, n2 = 0;
    bool binary = true;
    std::vector<int> rowptr, colidx;
    std::vector<double> val;   // empty when binary
    std::vector<double> eigs;  // eigenvalues of W (analytic on the torus, else user-supplied) -- or the nodes of a Lanczos
                               // quadrature of its spectral measure (spectrum.cu), then with weights:
    std::vector<double> eig_w; // empty = every node counts once (exact spectrum)
    double ew_min = 0, ew_max = 0;
    // transient, sorted by rho: sum log(1 - rho lam) for the distinct rho values of a surface call in progress
    // (mclcar_glm_surface) -- a grid of nr x ns candidates needs nr spectral sums, not nr * ns
    std::vector<std::pair<double, long double>> logdet_tab;
    int max_deg = 0;
    // greedy colouring: sites of colour c are order[color_ptr[c] .. color_ptr[c+1])
    int ncolors = 0;
    std::vector<int> color_ptr, order, color_of;
    int* d_rowptr = nullptr;
    int* d_colidx = nullptr;
    double* d_val = nullptr;
    int* d_order = nullptr;

    // ---- design
    bool has_design = false;
    int p = 0;
    std::vector<double> X, WX;   // n x p column-major
    std::vector<double> G0, G1;  // p x p: X'X, X'WX
    double* d_X = nullptr;       // [p][n] fp64
    double* d_WX = nullptr;

    // ---- observations
    bool has_obs = false;
    bool design_const = true;    // every design column is a constant (intercept-only models): X beta is one number
    // W X, X'X, X'WX (design) and the statistics of y (observations) are formed on first use: ensure_algebra()
    bool design_pending = false, obs_pending = false;
    void* h_wx = nullptr;        // pinned staging of W X on its way to the device
    size_t h_wx_bytes = 0;
    cudaEvent_t wx_event = nullptr;
    std::vector<double> y, ntrial;
    std::vector<double> qobs;  // (y'y, y'Wy, X'y[p], X'Wy[p])
    std::vector<double> h_scratch;  // W y of the last set_obs (host scratch)
    double* d_y = nullptr;
    double* d_ntrial = nullptr;

    // ---- scratch reused across calls (grown on demand)
    mclcar::DeviceBuf d_coef, d_partial, d_tuples, d_subidx, d_gather;
    mclcar::DeviceBuf d_sptab;   // piecewise-polynomial table of log1p(exp(-a)) for the binomial data term (glm.cu)
    void* h_pinned = nullptr;
    size_t h_pinned_bytes = 0;

    // ---- caching device allocator: big buffers (sample matrices, chain state, statistics) are
    // recycled across calls so that steady-state steps do no cudaMalloc / cudaFree
    struct PoolBlock { void* p; size_t bytes; bool used; };
    std::vector<PoolBlock> pool;
    size_t pool_cap = (size_t)64 << 30;  // cached free bytes allowed to linger (half the device memory)

    std::vector<mclcar::EdgeCache> edge_cache;

    // ---- hierarchical model (hcar.cu owns the type)
    void* hcar = nullptr;

    // ---- NCCL (dlopen'ed)
    void* nccl_lib = nullptr;
    void* nccl_comm = nullptr;
    // ---- in-process group (one R session driving several GPUs from host threads): the contexts of a group
    // exchange their partial tuples through host memory at the point where ranks of a job use NCCL
    void* host_group = nullptr;

    // ---- instrumentation: kernel launches per category and (when enabled) CUDA-event spans
    bool profiling = false;
    int64_t launches[mclcar::PROF_NCAT] = {0};
    struct Span { int cat; cudaEvent_t a, b; };
    std::vector<Span> spans;
    double span_ms[mclcar::PROF_NCAT] = {0};
};

struct mclcar_samples {
    mclcar_ctx* ctx = nullptr;
    int dtype = MCLCAR_F64;
    int layout = MCLCAR_COL_IS_SAMPLE;
    int family = 0;  // FAM_* of the attached model (GLM sample sets)
    int64_t sites = 0, N = 0, ld = 0;
    void* d_M = nullptr;  // may be null (statistics-only sample set)
    bool owned = false;
    // sufficient statistics, device, [d][N] fp64
    int d = 0;
    double* d_q = nullptr;
    bool stats_ready = false;
    std::vector<double> qbar;  // column means of q (host)
    // importance point
    bool attached = false;
    std::vector<double> psi0;
    double t10 = 0.0;
    double* d_base = nullptr;  // user-supplied t20 (N) or null => computed from psi0 and q
    double base_mean = 0.0;
    // sampler diagnostics
    double accept = 1.0;
    int64_t sweeps = 0;
    int64_t chains = 0;   // parallel chains the sampler ran (0: matrix supplied by the caller)
    int burn = 0, thin = 0;
    double omega = 0.0;
};

namespace mclcar {

int fail(mclcar_ctx* ctx, int code, const char* fmt, ...);
void set_global_error(const char* msg);

#define MCL_CUDA(ctx, expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return ::mclcar::fail((ctx), _e == cudaErrorMemoryAllocation ? MCLCAR_ENOMEM : MCLCAR_ECUDA, \
                                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define MCL_TRY(expr)                    \
    do {                                 \
        int _s = (expr);                 \
        if (_s != MCLCAR_OK) return _s;  \
    } while (0)

// RAII span: two events on the context's stream around a group of launches of one category
struct ProfScope {
    mclcar_ctx* ctx;
    int cat;
    cudaEvent_t a = nullptr;
    ProfScope(mclcar_ctx* c, int category, int64_t n_launches);
    void close();
    ~ProfScope() { close(); }
};

int pool_alloc(mclcar_ctx* ctx, void** out, size_t bytes);
void pool_free(mclcar_ctx* ctx, void* p);
void pool_release_all(mclcar_ctx* ctx);
void pool_trim(mclcar_ctx* ctx, size_t keep_bytes);
int ensure(mclcar_ctx* ctx, DeviceBuf& b, size_t bytes);
int ensure_pinned(mclcar_ctx* ctx, size_t bytes);
int require_device(mclcar_ctx* ctx);
int ensure_algebra(mclcar_ctx* ctx);   // call before reading WX, G0, G1, qobs or d_WX

// ---- statistics kernels (stats_torus.cu / stats_csr.cu) -----------------------------------
// q layout: [d][ldq]; d = 2 + 2p: (x'x, x'Wx, X'x[p], (WX)'x[p])
int launch_stats_site_major(mclcar_ctx* ctx, const void* M, int dtype, int64_t N, int64_t ld, double* d_q,
                            int64_t ldq);
int launch_transpose(mclcar_ctx* ctx, const void* in, int dtype, int64_t rows, int64_t cols, int64_t ld_in,
                     void* out, int64_t ld_out);
int launch_stats_torus_fast(mclcar_ctx* ctx, const void* M, int dtype, int64_t N, int64_t ld, double* d_q,
                            int64_t ldq, bool* handled);
int launch_quadforms_site_major(mclcar_ctx* ctx, const void* M, int dtype, int64_t N, int64_t ld, int n,
                                const std::vector<int>& rowptr, const std::vector<int>& colidx,
                                const std::vector<double>& val, int L, const double* const* d_vecs, double* d_out,
                                int64_t ldq);
int edge_cache_get(mclcar_ctx* ctx, int n, const std::vector<int>& rowptr, const std::vector<int>& colidx,
                   const std::vector<double>& val, const mclcar::EdgeCache** out);
int samples_alloc_stats(mclcar_ctx* ctx, mclcar_samples* s, int d);
int samples_update_qbar(mclcar_ctx* ctx, mclcar_samples* s);
int merge_tuples(int what, int F, int d, int G, int K, const double* in, double* out);
void centred_moments(const double* tup, int V, double* vbar, double* C);
int gather_and_merge(mclcar_ctx* ctx, int what, int F, int d, int K, double* tuples);
// NCCL all-gather of `cnt` doubles per rank, device to device, on the context's stream (no host round trip)
int nccl_all_gather_dev(mclcar_ctx* ctx, const double* d_send, double* d_recv, size_t cnt);
inline bool nccl_active(const mclcar_ctx* ctx) { return ctx->nccl_comm != nullptr && ctx->world > 1; }
void host_group_leave(mclcar_ctx* ctx);
void comm_destroy(mclcar_ctx* ctx);
void hcar_forget(mclcar_ctx* ctx);
void edge_cache_clear(mclcar_ctx* ctx);

// ---- reduction kernels (reduce.cu) ----------------------------------------------------------
struct ReducePlan {
    const double* d_q = nullptr;  // [d][ldq]
    int64_t ldq = 0;
    int d = 0;
    const double* d_base = nullptr;  // [N] or null
    int64_t N = 0;
    int K = 0;
    int F = 0;       // features with second moments (0 for lr/var)
    int what = WHAT_LR;
    int shift_mode = MCLCAR_SHIFT_MEAN;
    // host coefficient blocks, K x coef_stride (see coef_* helpers)
    std::vector<double> coef;
    int coef_stride = 0;
    // subsets (host, 0-based) or null
    const int64_t* subset_idx = nullptr;
    const int64_t* subset_ptr = nullptr;
};
// coefficient block of one candidate: c0, c[d], shift, n, centre[1+F], feat[F][1+d] (a0 then A)
inline int coef_stride_of(int d, int F) { return 3 + d + (1 + F) + F * (1 + d); }
inline int coef_off_shift(int d) { return 1 + d; }
inline int coef_off_n(int d) { return 2 + d; }
inline int coef_off_centre(int d) { return 3 + d; }
inline int coef_off_feat(int d, int F) { return 3 + d + (1 + F); }
int64_t tuple_len(int what, int F, int d);
// runs the reduction; leaves K tuples in ctx->d_tuples (device) and copies them to `tuples` (host).  With
// merge_ranks and an NCCL communicator the device tuples of all ranks are all-gathered on the stream first and ONE
// device-to-host copy brings everything back: one synchronisation per candidate batch; `tuples` then holds the
// rank-ordered merge (the same on every rank).
int run_reduce(mclcar_ctx* ctx, ReducePlan& plan, double* tuples, bool merge_ranks = false);


// ---- samplers (sampler_torus.cu / sampler_csr.cu) ---------------------------------------------
struct TorusChains {
    float* red = nullptr;    // current state
    float* black = nullptr;
    float* red2 = nullptr;   // ping-pong partner used by the fused sweep
    float* black2 = nullptr;
    int64_t C = 0;
};
int torus_chains_alloc(mclcar_ctx* ctx, int64_t C, TorusChains& t);
void torus_chains_free(mclcar_ctx* ctx, TorusChains& t);
struct TorusEmit {   // where the black half sweep of an emitting sweep writes the samples
    void* out = nullptr;
    int dtype = MCLCAR_F32;
    int64_t ld = 0, slot0 = 0, nch = 0;
    const double* d_mu = nullptr;  // per-site mean (device) or null => mu_const
    double mu_const = 0.0;
    // statistics-only emission: the sweep accumulates the sufficient statistics of the samples it would have
    // written into q[j * ldq + slot0 + chain] instead of writing them (constant mean and design columns only)
    double* stats_q = nullptr;
    int64_t stats_ldq = 0;
};
int torus_fused_tile_rows(const mclcar_ctx* ctx);
int torus_sweep(mclcar_ctx* ctx, TorusChains& t, double rho, double sigma, double omega, uint32_t sweep,
                const TorusEmit* emit = nullptr, int want = 1, int* done = nullptr);
int torus_emit(mclcar_ctx* ctx, TorusChains& t, const double* d_mu, double mu_const, void* out, int dtype, int64_t ld,
               int64_t slot0, int64_t nch);

enum { FAM_GAUSS = 0, FAM_BINOM = 1, FAM_POIS = 2 };
// host description (site order, fp64) of a Gaussian Markov random field exp(-x'Qx/2 + b'x)
struct GmrfHost {
    int n = 0;
    const std::vector<int>* rowptr = nullptr;  // CSR of the off-diagonal pattern of Q
    const std::vector<int>* colidx = nullptr;
    std::vector<double> qoff;   // Q_sj per CSR entry
    std::vector<double> qdiag;  // Q_ss
    std::vector<double> b;      // linear term (empty = 0)
    const std::vector<int>* color_ptr = nullptr;  // colouring: positions -> sites
    const std::vector<int>* order = nullptr;
};
int run_csr_sampler(mclcar_ctx* ctx, const GmrfHost& g, int fam, double omega, const double* mu_host, const double* y,
                    const double* ntrial, const double* xb, int64_t N, int64_t C_want, int burn, int thin,
                    void* d_out, int out_f64, int64_t ld, double* accept_rate, int64_t* sweeps_out, double* d_stat_q = nullptr,
                    int64_t stat_ldq = 0, bool* stats_done = nullptr);
int check_rho(mclcar_ctx* ctx, double rho);
// sum_i w_i f(lambda_i) over the context's spectral nodes: tr f(W), exactly (eigenvalues) or by quadrature
template <class Fn>
inline long double spectral_sum(const mclcar_ctx* ctx, Fn f) {
    long double a = 0;
    if (ctx->eig_w.empty()) for (double l : ctx->eigs) a += f(l);
    else for (size_t i = 0; i < ctx->eigs.size(); ++i) a += (long double)ctx->eig_w[i] * f(ctx->eigs[i]);
    return a;
}
// make sure the context has a spectrum: estimates one on the device when the caller supplied none
int ensure_spectrum(mclcar_ctx* ctx);
// sum_i w_i log(1 - rho_k lambda_i) for K values of rho (exact.cu)
int logdet_sums(mclcar_ctx* ctx, const double* rho, int K, double* out);
// start of a sampler run: derive this call's Philox key and advance the context's draw counter
inline uint64_t begin_draw(mclcar_ctx* ctx) {
    ctx->draw_key = ctx->seed + ctx->draw * 0x9E3779B97F4A7C15ull + ctx->draw_salt * 0xD1B54A32D192ED03ull;
    ++ctx->draw;
    return ctx->draw_key;
}
void default_schedule(const mclcar_ctx* ctx, double rho, double* omega, int* burn, int* thin);
int64_t default_csr_chains(const mclcar_ctx* ctx, int n, int64_t N, int burn = 0, int thin = 0);
int64_t default_torus_chains(int n, int64_t N, int burn, int thin);
int dcar_ensure_stats(mclcar_ctx* ctx, mclcar_samples* s);

}  // namespace mclcar

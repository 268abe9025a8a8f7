// K3 -- importance-weight reduction for a whole batch of candidate parameters.
//
// For candidate k and sample i the log weight is affine in the sample's sufficient
// statistics q_i ("sufficient-statistic restatement"):
//     l_ik = c0_k + sum_j c_kj q_ij - base_i,        e_ik = exp(l_ik - m_k)
// and every per-sample derivative the reference forms (D.log.unorm, R/mcl.dCAR.R:188-210;
// grad.logH, R/mcl.bin.R:188-196; grad.uy, R/mcl.HCAR.R:187-196) is an affine feature
//     f_ika = a0_ka + sum_j A_kaj q_ij.
// The kernels accumulate, per candidate, the partial tuples from which the host forms
// exactly the reference's estimators (log-mean-exp, unbiased weight variance, weighted
// gradient mean, gradient covariance, Hessian parts):
//   MOMENTS: with v = (e, e f_1 .. e f_F) and a fixed centre c:  sum (v - c),  sum (v - c)(v - c)'
//   HESS   : sum e, sum e f, sum e f f'   and   sum e q_j  (first moments of the raw statistics)
// The shift m_k and the centre come with the coefficient block (mean shift: computed on
// the host from the column means of q; max shift: k3_prepare_max below).  Centred sums
// keep the one-pass variance free of cancellation; tuples of different shards / GPUs are
// merged by exact re-centring on the host (tuples.cpp).
//
// Work decomposition: grid = (sample splits S, candidate tiles); each thread strides over
// samples, holds the accumulators of KT candidates in registers, warp-shuffle + shared
// memory tree per CTA, one partial per (candidate, split); k3_finalize adds the S
// partials in fixed order => deterministic results.
#include "common.cuh"
#include "engine.hpp"

namespace mclcar {

namespace {

constexpr int kThreads = 256;

struct K3Params {
    const double* q;
    int64_t ldq;
    int d;
    const double* base;
    int64_t N;
    const double* coef;  // [K][cs]
    int cs;
    int K;
    const int32_t* sub_idx;  // device pool or null
    const int64_t* sub_ptr;  // device [K+1] or null
    double* partial;         // [K][S][T]
    int S;
    int T;
};

__device__ __forceinline__ int64_t subset_range(const K3Params& p, int k, int64_t& off) {
    if (p.sub_ptr != nullptr) {
        const int64_t a = p.sub_ptr[k], b = p.sub_ptr[k + 1];
        if (b > a) {
            off = a;
            return b - a;
        }
    }
    off = -1;
    return p.N;
}

// ---------------------------------------------------------------------------------------
// max shift + pilot centre (MCLCAR_SHIFT_MAX only).  One CTA per candidate.
// ---------------------------------------------------------------------------------------
template <int FMAX>
__global__ void __launch_bounds__(kThreads) k3_prepare_max(K3Params p, double* coef_rw, int F) {
    extern __shared__ double sm[];
    const int k = blockIdx.x;
    const int d = p.d;
    double* cf = sm;  // coefficient block
    for (int t = threadIdx.x; t < p.cs; t += blockDim.x) cf[t] = p.coef[(size_t)k * p.cs + t];
    __syncthreads();
    int64_t off;
    const int64_t ns = subset_range(p, k, off);
    double mx = -INFINITY;
    for (int64_t t = threadIdx.x; t < ns; t += blockDim.x) {
        const int64_t i = off >= 0 ? (int64_t)p.sub_idx[off + t] : t;
        double l = cf[0];
        for (int j = 0; j < d; ++j) l = fma(cf[1 + j], p.q[(size_t)j * p.ldq + i], l);
        if (p.base) l -= p.base[i];
        mx = fmax(mx, l);
    }
    __shared__ double red[kThreads / 32];
    __shared__ double s_m;
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = red[0];
        for (int w = 1; w < kThreads / 32; ++w) a = fmax(a, red[w]);
        s_m = a;
    }
    __syncthreads();
    const double m = s_m;
    // pilot centre: mean of v over the first <= 1024 samples
    const int64_t npilot = ns < 1024 ? ns : 1024;
    double acc[1 + FMAX];
#pragma unroll
    for (int a = 0; a <= FMAX; ++a) acc[a] = 0.0;
    const double* feat = cf + 3 + d + (1 + F);
    for (int64_t t = threadIdx.x; t < npilot; t += blockDim.x) {
        const int64_t i = off >= 0 ? (int64_t)p.sub_idx[off + t] : t;
        double l = cf[0];
        for (int j = 0; j < d; ++j) l = fma(cf[1 + j], p.q[(size_t)j * p.ldq + i], l);
        if (p.base) l -= p.base[i];
        const double e = exp(l - m);
        acc[0] += e;
#pragma unroll
        for (int a = 0; a < FMAX; ++a) {
            if (a < F) {
                double f = feat[a * (1 + d)];
                for (int j = 0; j < d; ++j) f = fma(feat[a * (1 + d) + 1 + j], p.q[(size_t)j * p.ldq + i], f);
                acc[1 + a] += e * f;
            }
        }
    }
    __shared__ double scratch[(kThreads / 32) * (1 + FMAX)];
    block_sum<1 + FMAX>(acc, scratch);
    if (threadIdx.x == 0) {
        double* out = coef_rw + (size_t)k * p.cs;
        out[1 + d] = m;
        for (int a = 0; a <= F; ++a) out[3 + d + a] = npilot > 0 ? acc[a] / (double)npilot : 0.0;
    }
}

// ---------------------------------------------------------------------------------------
// centred first and second moments of v = (e, e f).  F == 0 is the lr / var fast path.
// ---------------------------------------------------------------------------------------
template <int F, int KT>
__global__ void __launch_bounds__(kThreads) k3_moments(K3Params p) {
    constexpr int V = 1 + F;
    constexpr int NS2 = V * (V + 1) / 2;
    constexpr int NV = V + NS2;
    extern __shared__ double sm[];
    const int d = p.d;
    const int k0 = blockIdx.y * KT;
    double* cf = sm;  // KT coefficient blocks
    for (int t = threadIdx.x; t < KT * p.cs; t += blockDim.x) {
        const int kk = t / p.cs;
        cf[t] = (k0 + kk < p.K) ? p.coef[(size_t)(k0 + kk) * p.cs + (t - kk * p.cs)] : 0.0;
    }
    __syncthreads();
    int64_t off;
    const int64_t ns = subset_range(p, k0, off);  // KT > 1 is only launched without subsets

    double acc[KT][NV];
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
#pragma unroll
        for (int a = 0; a < NV; ++a) acc[kk][a] = 0.0;

    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ns; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = off >= 0 ? (int64_t)p.sub_idx[off + t] : t;
        double l[KT], f[KT][F > 0 ? F : 1];
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
            l[kk] = cf[kk * p.cs];
#pragma unroll
            for (int a = 0; a < F; ++a) f[kk][a] = cf[kk * p.cs + 3 + d + V + a * (1 + d)];
        }
        for (int j = 0; j < d; ++j) {
            const double qj = p.q[(size_t)j * p.ldq + i];
#pragma unroll
            for (int kk = 0; kk < KT; ++kk) {
                l[kk] = fma(cf[kk * p.cs + 1 + j], qj, l[kk]);
#pragma unroll
                for (int a = 0; a < F; ++a) f[kk][a] = fma(cf[kk * p.cs + 3 + d + V + a * (1 + d) + 1 + j], qj, f[kk][a]);
            }
        }
        const double b = p.base ? p.base[i] : 0.0;
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
            const double* c = cf + kk * p.cs;
            const double e = exp(l[kk] - b - c[1 + d]);
            double dv[V];
            dv[0] = e - c[3 + d];
#pragma unroll
            for (int a = 0; a < F; ++a) dv[1 + a] = fma(e, f[kk][a], -c[3 + d + 1 + a]);
            int s = V;
#pragma unroll
            for (int a = 0; a < V; ++a) {
                acc[kk][a] += dv[a];
#pragma unroll
                for (int bb = a; bb < V; ++bb) {
                    acc[kk][s] = fma(dv[a], dv[bb], acc[kk][s]);
                    ++s;
                }
            }
        }
    }
    // CTA reduction, one candidate at a time
    double* scratch = sm + KT * p.cs;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
        block_sum<NV>(acc[kk], scratch);
        if (threadIdx.x == 0 && k0 + kk < p.K) {
            double* out = p.partial + ((size_t)(k0 + kk) * p.S + blockIdx.x) * p.T + 2 + V;
#pragma unroll
            for (int a = 0; a < NV; ++a) out[a] = acc[kk][a];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Hessian parts: sum e, sum e f, sum e f f' (uncentred, as the reference forms them).
// ---------------------------------------------------------------------------------------
template <int F>
__global__ void __launch_bounds__(kThreads) k3_hess(K3Params p) {
    constexpr int NV = 1 + F + F * (F + 1) / 2;
    extern __shared__ double sm[];
    const int d = p.d;
    const int k = blockIdx.y;
    double* cf = sm;
    for (int t = threadIdx.x; t < p.cs; t += blockDim.x) cf[t] = p.coef[(size_t)k * p.cs + t];
    __syncthreads();
    int64_t off;
    const int64_t ns = subset_range(p, k, off);
    double acc[NV];
#pragma unroll
    for (int a = 0; a < NV; ++a) acc[a] = 0.0;
    const double* feat = cf + 3 + d + (1 + F);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ns; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = off >= 0 ? (int64_t)p.sub_idx[off + t] : t;
        double l = cf[0], f[F > 0 ? F : 1];
#pragma unroll
        for (int a = 0; a < F; ++a) f[a] = feat[a * (1 + d)];
        for (int j = 0; j < d; ++j) {
            const double qj = p.q[(size_t)j * p.ldq + i];
            l = fma(cf[1 + j], qj, l);
#pragma unroll
            for (int a = 0; a < F; ++a) f[a] = fma(feat[a * (1 + d) + 1 + j], qj, f[a]);
        }
        if (p.base) l -= p.base[i];
        const double e = exp(l - cf[1 + d]);
        acc[0] += e;
        int s = 1 + F;
#pragma unroll
        for (int a = 0; a < F; ++a) {
            const double ef = e * f[a];
            acc[1 + a] += ef;
#pragma unroll
            for (int bb = a; bb < F; ++bb) {
                acc[s] = fma(ef, f[bb], acc[s]);
                ++s;
            }
        }
    }
    double* scratch = sm + p.cs;
    block_sum<NV>(acc, scratch);
    if (threadIdx.x == 0) {
        double* out = p.partial + ((size_t)k * p.S + blockIdx.x) * p.T + 2;
#pragma unroll
        for (int a = 0; a < NV; ++a) out[a] = acc[a];
    }
}

// first moments of the raw statistics: sum_i e_ik q_ij for j in [j0, j0 + nr), nr <= 16
__global__ void __launch_bounds__(kThreads) k3_first(K3Params p, int j0, int nr, int out_off) {
    constexpr int NR = 16;
    extern __shared__ double sm[];
    const int d = p.d;
    const int k = blockIdx.y;
    double* cf = sm;
    for (int t = threadIdx.x; t < p.cs; t += blockDim.x) cf[t] = p.coef[(size_t)k * p.cs + t];
    __syncthreads();
    int64_t off;
    const int64_t ns = subset_range(p, k, off);
    double acc[NR];
#pragma unroll
    for (int a = 0; a < NR; ++a) acc[a] = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ns; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = off >= 0 ? (int64_t)p.sub_idx[off + t] : t;
        double l = cf[0];
        for (int j = 0; j < d; ++j) l = fma(cf[1 + j], p.q[(size_t)j * p.ldq + i], l);
        if (p.base) l -= p.base[i];
        const double e = exp(l - cf[1 + d]);
#pragma unroll
        for (int a = 0; a < NR; ++a)
            if (a < nr) acc[a] = fma(e, p.q[(size_t)(j0 + a) * p.ldq + i], acc[a]);
    }
    double* scratch = sm + p.cs;
    block_sum<NR>(acc, scratch);
    if (threadIdx.x == 0) {
        double* out = p.partial + ((size_t)k * p.S + blockIdx.x) * p.T + out_off + j0;
        for (int a = 0; a < nr; ++a) out[a] = acc[a];
    }
}

// add the S partials of every candidate in split order; fill n, shift and centre
__global__ void k3_finalize(K3Params p, const double* coef, double* tuples, int n_copy /* centre entries */) {
    const int k = blockIdx.x;
    const int d = p.d;
    for (int t = threadIdx.x; t < p.T; t += blockDim.x) {
        double v;
        if (t == 0) {
            v = coef[(size_t)k * p.cs + 2 + d];
        } else if (t == 1) {
            v = coef[(size_t)k * p.cs + 1 + d];
        } else if (t < 2 + n_copy) {
            v = coef[(size_t)k * p.cs + 3 + d + (t - 2)];
        } else {
            v = 0.0;
            for (int s = 0; s < p.S; ++s) v += p.partial[((size_t)k * p.S + s) * p.T + t];
        }
        tuples[(size_t)k * p.T + t] = v;
    }
}

template <int F>
void launch_moments(const K3Params& p, dim3 grid_kt1, dim3 grid_kt4, bool tile4, size_t scratch_doubles,
                    cudaStream_t st) {
    if constexpr (F == 0) {
        if (tile4) {
            const size_t sh = (4 * (size_t)p.cs + scratch_doubles) * sizeof(double);
            k3_moments<0, 4><<<grid_kt4, kThreads, sh, st>>>(p);
            return;
        }
    }
    const size_t sh = ((size_t)p.cs + scratch_doubles) * sizeof(double);
    k3_moments<F, 1><<<grid_kt1, kThreads, sh, st>>>(p);
}

}  // namespace

// every K3 instantiation may use up to 200 KB of dynamic shared memory (set once per process)
static int k3_allow_large_smem(mclcar_ctx* ctx) {
    static bool done = false;
    if (done) return MCLCAR_OK;
    const int big = 200 * 1024;
#define ALLOW(fn) MCL_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, big))
    ALLOW((k3_prepare_max<MCLCAR_MAX_P>));
    ALLOW((k3_moments<0, 4>)); ALLOW((k3_moments<0, 1>)); ALLOW((k3_moments<2, 1>)); ALLOW((k3_moments<4, 1>));
    ALLOW((k3_moments<6, 1>)); ALLOW((k3_moments<8, 1>)); ALLOW((k3_moments<10, 1>)); ALLOW((k3_moments<12, 1>));
    ALLOW(k3_hess<2>); ALLOW(k3_hess<4>); ALLOW(k3_hess<6>); ALLOW(k3_hess<8>); ALLOW(k3_hess<10>); ALLOW(k3_hess<12>);
    ALLOW(k3_first);
#undef ALLOW
    done = true;
    return MCLCAR_OK;
}

int64_t tuple_len(int what, int F, int d) {
    const int V = 1 + F;
    switch (what) {
        case WHAT_LR: return 5;
        case WHAT_GRAD: return 2 + 2 * V + V * (V + 1) / 2;
        case WHAT_HESS: return 3 + F + F * (F + 1) / 2 + d;
        default: return -1;
    }
}

static int run_reduce_body(mclcar_ctx* ctx, ReducePlan& plan, double* tuples_host, bool merge_ranks, bool* gathered);

// A rank that fails on its own (a CUDA error, an allocation) before it reaches the tuple exchange must not leave the
// other ranks waiting in ncclAllGather: it still takes part, with tuples of all-ones bytes (NaN), and every rank --
// this one included -- returns an error for the batch.  (Argument errors are the same on every rank: the candidates
// and the model are replicated, so those ranks all stop before the exchange.)
int run_reduce(mclcar_ctx* ctx, ReducePlan& plan, double* tuples_host, bool merge_ranks) {
    bool gathered = false;
    const int st = run_reduce_body(ctx, plan, tuples_host, merge_ranks, &gathered);
    if (st != MCLCAR_OK && merge_ranks && nccl_active(ctx) && !gathered) {
        const std::string keep = ctx->err;
        const int F = plan.what == WHAT_LR ? 0 : plan.F;
        const size_t cnt = (size_t)plan.K * (size_t)tuple_len(plan.what, (F + 1) & ~1, plan.d);
        if (ensure(ctx, ctx->d_tuples, cnt * sizeof(double)) == MCLCAR_OK &&
            ensure(ctx, ctx->d_gather, (size_t)ctx->world * cnt * sizeof(double)) == MCLCAR_OK &&
            cudaMemsetAsync(ctx->d_tuples.p, 0xFF, cnt * sizeof(double), ctx->stream) == cudaSuccess &&
            nccl_all_gather_dev(ctx, (const double*)ctx->d_tuples.p, (double*)ctx->d_gather.p, cnt) == MCLCAR_OK)
            cudaStreamSynchronize(ctx->stream);
        ctx->err = keep;
    }
    return st;
}

static int run_reduce_body(mclcar_ctx* ctx, ReducePlan& plan, double* tuples_host, bool merge_ranks, bool* gathered) {
    MCL_TRY(require_device(ctx));
    const int K = plan.K, d = plan.d;
    const int F = plan.what == WHAT_LR ? 0 : plan.F;
    if (F > MCLCAR_MAX_P) return fail(ctx, MCLCAR_ELIMIT, "parameter vector longer than MCLCAR_MAX_P");
    const int Fr = (F + 1) & ~1;  // kernels are instantiated for even F; pad with zero features
    const int cs = plan.coef_stride;
    const int64_t T = tuple_len(plan.what, F, d);
    cudaStream_t st = ctx->stream;

    // re-pack the coefficient blocks for the padded feature count
    const int csr = coef_stride_of(d, Fr);
    std::vector<double> packed((size_t)K * csr, 0.0);
    for (int k = 0; k < K; ++k) {
        const double* src = plan.coef.data() + (size_t)k * cs;
        double* dst = packed.data() + (size_t)k * csr;
        // host layout (see eval_*.cpp): c0, c[d], shift, n, centre[1+F], feat[F][1+d]
        for (int t = 0; t < 3 + d; ++t) dst[t] = src[t];
        for (int a = 0; a <= F; ++a) dst[3 + d + a] = src[3 + d + a];
        for (int a = 0; a < F; ++a)
            for (int j = 0; j <= d; ++j) dst[3 + d + (1 + Fr) + a * (1 + d) + j] = src[3 + d + (1 + F) + a * (1 + d) + j];
    }
    // padded tuple length used on the device
    const int64_t Tr = tuple_len(plan.what, Fr, d);

    // subsets -> device
    const bool has_sub = plan.subset_ptr != nullptr;
    int32_t* d_idx = nullptr;
    int64_t* d_ptr = nullptr;
    int64_t max_ns = plan.N;
    if (has_sub) {
        const int64_t tot = plan.subset_ptr[K];
        std::vector<int32_t> idx32((size_t)(tot > 0 ? tot : 1));
        for (int64_t t = 0; t < tot; ++t) {
            if (plan.subset_idx[t] < 0 || plan.subset_idx[t] >= plan.N)
                return fail(ctx, MCLCAR_EINVAL, "subset index %lld out of range", (long long)plan.subset_idx[t]);
            idx32[t] = (int32_t)plan.subset_idx[t];
        }
        const size_t bytes_idx = idx32.size() * sizeof(int32_t), bytes_ptr = (size_t)(K + 1) * sizeof(int64_t);
        const size_t off_ptr = (bytes_idx + 15) & ~(size_t)15;
        MCL_TRY(ensure(ctx, ctx->d_subidx, off_ptr + bytes_ptr));
        d_idx = (int32_t*)ctx->d_subidx.p;
        d_ptr = (int64_t*)((char*)ctx->d_subidx.p + off_ptr);
        MCL_CUDA(ctx, cudaMemcpyAsync(d_idx, idx32.data(), bytes_idx, cudaMemcpyHostToDevice, st));
        MCL_CUDA(ctx, cudaMemcpyAsync(d_ptr, plan.subset_ptr, bytes_ptr, cudaMemcpyHostToDevice, st));
        MCL_CUDA(ctx, cudaStreamSynchronize(st));  // idx32 is a local
    }

    // launch geometry
    const bool tile4 = (F == 0) && !has_sub && K >= 4;
    const int ktiles = tile4 ? (K + 3) / 4 : K;
    int S = (4 * ctx->n_sm + ktiles - 1) / ktiles;
    const int64_t max_splits = (max_ns + kThreads - 1) / kThreads;
    if (S > max_splits) S = (int)max_splits;
    if (S < 1) S = 1;

    MCL_TRY(ensure(ctx, ctx->d_coef, packed.size() * sizeof(double)));
    MCL_TRY(ensure(ctx, ctx->d_partial, (size_t)K * S * Tr * sizeof(double)));
    MCL_TRY(ensure(ctx, ctx->d_tuples, (size_t)K * Tr * sizeof(double)));
    MCL_CUDA(ctx, cudaMemcpyAsync(ctx->d_coef.p, packed.data(), packed.size() * sizeof(double), cudaMemcpyHostToDevice, st));

    K3Params p;
    p.q = plan.d_q; p.ldq = plan.ldq; p.d = d; p.base = plan.d_base; p.N = plan.N;
    p.coef = (const double*)ctx->d_coef.p; p.cs = csr; p.K = K;
    p.sub_idx = d_idx; p.sub_ptr = d_ptr;
    p.partial = (double*)ctx->d_partial.p; p.S = S; p.T = (int)Tr;

    // the kernels keep a candidate's coefficient block (x 4 for the lr tile) in dynamic shared memory: d grows with
    // the number of distinct betas of a GLM batch -- opt in beyond 48 KB, refuse beyond what the SM has
    {
        const size_t need = ((size_t)(tile4 ? 4 : 1) * csr + (kThreads / 32) * (size_t)(1 + Fr + (1 + Fr) * (2 + Fr) / 2 + 16)) * sizeof(double);
        if (need > (size_t)200 * 1024)
            return fail(ctx, MCLCAR_ELIMIT, "candidate batch too wide for the reduction kernels: %zu bytes of coefficients per candidate "
                                           "(split the batch: fewer distinct betas per call)", (size_t)csr * sizeof(double));
        if (need > (size_t)40 * 1024) MCL_TRY(k3_allow_large_smem(ctx));
    }
    ProfScope prof(ctx, PROF_REDUCE, 2 + (plan.shift_mode == MCLCAR_SHIFT_MAX ? 1 : 0) +
                                         (plan.what == WHAT_HESS ? (d + 15) / 16 : 0));
    if (plan.shift_mode == MCLCAR_SHIFT_MAX) {
        k3_prepare_max<MCLCAR_MAX_P><<<K, kThreads, (size_t)csr * sizeof(double), st>>>(p, (double*)ctx->d_coef.p, Fr);
    }

    const size_t warps = kThreads / 32;
    const dim3 g1(S, K), g4(S, (K + 3) / 4);
    if (plan.what == WHAT_HESS) {
        const size_t nv = 1 + Fr + Fr * (Fr + 1) / 2;
        const size_t sh = ((size_t)csr + warps * nv) * sizeof(double);
        switch (Fr) {
            case 2: k3_hess<2><<<g1, kThreads, sh, st>>>(p); break;
            case 4: k3_hess<4><<<g1, kThreads, sh, st>>>(p); break;
            case 6: k3_hess<6><<<g1, kThreads, sh, st>>>(p); break;
            case 8: k3_hess<8><<<g1, kThreads, sh, st>>>(p); break;
            case 10: k3_hess<10><<<g1, kThreads, sh, st>>>(p); break;
            case 12: k3_hess<12><<<g1, kThreads, sh, st>>>(p); break;
            default: return fail(ctx, MCLCAR_ELIMIT, "unsupported feature count %d", Fr);
        }
        const int out_off = 3 + Fr + Fr * (Fr + 1) / 2;
        const size_t sh1 = ((size_t)csr + warps * 16) * sizeof(double);
        for (int j0 = 0; j0 < d; j0 += 16) {
            const int nr = d - j0 < 16 ? d - j0 : 16;
            k3_first<<<g1, kThreads, sh1, st>>>(p, j0, nr, out_off);
        }
    } else {
        const size_t V = 1 + Fr, nv = V + V * (V + 1) / 2;
        const size_t scratch = warps * nv;
        switch (Fr) {
            case 0: launch_moments<0>(p, g1, g4, tile4, scratch, st); break;
            case 2: launch_moments<2>(p, g1, g4, false, scratch, st); break;
            case 4: launch_moments<4>(p, g1, g4, false, scratch, st); break;
            case 6: launch_moments<6>(p, g1, g4, false, scratch, st); break;
            case 8: launch_moments<8>(p, g1, g4, false, scratch, st); break;
            case 10: launch_moments<10>(p, g1, g4, false, scratch, st); break;
            case 12: launch_moments<12>(p, g1, g4, false, scratch, st); break;
            default: return fail(ctx, MCLCAR_ELIMIT, "unsupported feature count %d", Fr);
        }
    }
    const int n_copy = plan.what == WHAT_HESS ? 0 : 1 + Fr;
    k3_finalize<<<K, 128, 0, st>>>(p, (const double*)ctx->d_coef.p, (double*)ctx->d_tuples.p, n_copy);
    MCL_CUDA(ctx, cudaGetLastError());
    prof.close();

    // device tuples (padded F) -> host tuples (exact F); with several ranks: all-gather on the device first, then
    // ONE copy and ONE synchronisation for the whole candidate batch, and the rank-ordered merge on the host
    const bool merge = merge_ranks && nccl_active(ctx);
    const int G = merge ? ctx->world : 1;
    const size_t cnt = (size_t)K * Tr;
    const double* d_src = (const double*)ctx->d_tuples.p;
    if (merge) {
        MCL_TRY(ensure(ctx, ctx->d_gather, (size_t)G * cnt * sizeof(double)));
        *gathered = true;
        MCL_TRY(nccl_all_gather_dev(ctx, (const double*)ctx->d_tuples.p, (double*)ctx->d_gather.p, cnt));
        d_src = (const double*)ctx->d_gather.p;
    }
    MCL_TRY(ensure_pinned(ctx, (size_t)G * cnt * sizeof(double)));
    double* hp = (double*)ctx->h_pinned;
    MCL_CUDA(ctx, cudaMemcpyAsync(hp, d_src, (size_t)G * cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
    MCL_CUDA(ctx, cudaStreamSynchronize(st));
    const int V = 1 + F, Vr = 1 + Fr;
    std::vector<double> all;
    if (merge) all.resize((size_t)G * K * T);
    for (int g = 0; g < G; ++g)
        for (int k = 0; k < K; ++k) {
            const double* src = hp + (size_t)g * cnt + (size_t)k * Tr;
            double* dst = (merge ? all.data() + (size_t)g * K * T : tuples_host) + (size_t)k * T;
            dst[0] = src[0];
            dst[1] = src[1];
            if (plan.what == WHAT_HESS) {
                for (int a = 0; a <= F; ++a) dst[2 + a] = src[2 + a];
                int s = 0, sr = 0;
                for (int a = 0; a < Fr; ++a)
                    for (int b = a; b < Fr; ++b, ++sr)
                        if (a < F && b < F) dst[3 + F + s++] = src[3 + Fr + sr];
                for (int j = 0; j < d; ++j) dst[3 + F + F * (F + 1) / 2 + j] = src[3 + Fr + Fr * (Fr + 1) / 2 + j];
            } else {
                for (int a = 0; a < V; ++a) {
                    dst[2 + a] = src[2 + a];               // centre
                    dst[2 + V + a] = src[2 + Vr + a];      // S1
                }
                int s = 0, sr = 0;
                for (int a = 0; a < Vr; ++a)
                    for (int b = a; b < Vr; ++b, ++sr)
                        if (a < V && b < V) dst[2 + 2 * V + s++] = src[2 + 2 * Vr + sr];
            }
        }
    if (merge) {
        for (int g = 0; g < G; ++g)
            if (!(all[(size_t)g * K * T] >= 0.0))   // a rank that failed before the collective sends NaN tuples
                return fail(ctx, MCLCAR_ENCCL, "rank %d failed before the tuple exchange", g);
        MCL_TRY(merge_tuples(plan.what, F, d, G, K, all.data(), tuples_host));
    }
    return MCLCAR_OK;
}

}  // namespace mclcar

// Sample-set handles: the reference's simY / Zy / u.y matrices on the device.
#include <cstring>

#include <cmath>

#include "common.cuh"
#include "engine.hpp"

namespace mclcar {

namespace {

// column means of q [d][ldq] over the first N entries of each row; one CTA per row
__global__ void rowmean_kernel(const double* q, int64_t ldq, int64_t N, double* out) {
    __shared__ double scratch[8];
    double a[1] = {0.0};
    const double* r = q + (size_t)blockIdx.x * ldq;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) a[0] += r[i];
    block_sum<1>(a, scratch);
    if (threadIdx.x == 0) out[blockIdx.x] = a[0] / (double)N;
}

template <typename T>
__global__ void to_f64_kernel(const T* in, double* out, int64_t rows, int64_t cols, int64_t ld_in, int64_t ld_out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (c < cols && r < rows) out[r * ld_out + c] = (double)in[r * ld_in + c];
}

}  // namespace

int samples_alloc_stats(mclcar_ctx* ctx, mclcar_samples* s, int d) {
    if (s->d_q && s->d == d) return MCLCAR_OK;
    if (s->d_q) {
        pool_free(ctx, s->d_q);
        s->d_q = nullptr;
    }
    s->d = d;
    MCL_TRY(pool_alloc(ctx, (void**)&s->d_q, (size_t)d * s->N * sizeof(double)));
    s->stats_ready = false;
    return MCLCAR_OK;
}

int samples_update_qbar(mclcar_ctx* ctx, mclcar_samples* s) {
    MCL_TRY(ensure(ctx, ctx->d_tuples, (size_t)s->d * sizeof(double)));
    ProfScope prof(ctx, PROF_STATS, 1);
    rowmean_kernel<<<s->d, 256, 0, ctx->stream>>>(s->d_q, s->N, s->N, (double*)ctx->d_tuples.p);
    MCL_CUDA(ctx, cudaGetLastError());
    s->qbar.assign(s->d, 0.0);
    MCL_CUDA(ctx, cudaMemcpyAsync(s->qbar.data(), ctx->d_tuples.p, (size_t)s->d * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCLCAR_OK;
}

}  // namespace mclcar

using namespace mclcar;

static int make_samples(mclcar_ctx* ctx, int dtype, int layout, int64_t sites, int64_t N, int64_t ld,
                        mclcar_samples** out) {
    if (!ctx || !out) return MCLCAR_EINVAL;
    *out = nullptr;
    if (dtype != MCLCAR_F64 && dtype != MCLCAR_F32) return fail(ctx, MCLCAR_EINVAL, "bad dtype %d", dtype);
    if (layout != MCLCAR_COL_IS_SAMPLE && layout != MCLCAR_ROW_IS_SAMPLE) return fail(ctx, MCLCAR_EINVAL, "bad layout %d", layout);
    if (sites < 1 || N < 1) return fail(ctx, MCLCAR_EINVAL, "empty sample matrix (%lld sites x %lld samples)", (long long)sites, (long long)N);
    if (N >= ((int64_t)1 << 31)) return fail(ctx, MCLCAR_ELIMIT, "more than 2^31 samples on one context");
    const int64_t lead = layout == MCLCAR_COL_IS_SAMPLE ? sites : N;
    if (ld < lead) return fail(ctx, MCLCAR_EINVAL, "ld %lld smaller than the leading extent %lld", (long long)ld, (long long)lead);
    mclcar_samples* s = new (std::nothrow) mclcar_samples();
    if (!s) return fail(ctx, MCLCAR_ENOMEM, "out of host memory");
    s->ctx = ctx; s->dtype = dtype; s->layout = layout; s->sites = sites; s->N = N; s->ld = ld;
    *out = s;
    return MCLCAR_OK;
}

extern "C" {

int mclcar_samples_from_host(mclcar_ctx* ctx, const void* M, int dtype, int layout, int64_t sites, int64_t N,
                             int64_t ld, mclcar_samples** out) {
    if (!ctx || !M) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    mclcar_samples* s = nullptr;
    MCL_TRY(make_samples(ctx, dtype, layout, sites, N, ld, &s));
    const size_t es = dtype == MCLCAR_F32 ? 4 : 8;
    const int64_t outer = layout == MCLCAR_COL_IS_SAMPLE ? N : sites;
    const int64_t lead = layout == MCLCAR_COL_IS_SAMPLE ? sites : N;
    // device copy is packed to a 16-byte aligned leading dimension (bulk-copy requirement)
    const int64_t align = 16 / es;
    const int64_t ldd = (lead + align - 1) / align * align;
    int stp = pool_alloc(ctx, &s->d_M, (size_t)outer * ldd * es);
    if (stp != MCLCAR_OK) {
        delete s;
        return stp;
    }
    s->owned = true;
    s->ld = ldd;
    cudaError_t e = cudaMemcpy2DAsync(s->d_M, (size_t)ldd * es, M, (size_t)ld * es, (size_t)lead * es, (size_t)outer,
                          cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        mclcar_samples_free(s);
        return fail(ctx, MCLCAR_ECUDA, "host to device copy of the sample matrix failed: %s", cudaGetErrorString(e));
    }
    *out = s;
    return MCLCAR_OK;
}

int mclcar_samples_from_device(mclcar_ctx* ctx, const void* M_dev, int dtype, int layout, int64_t sites, int64_t N,
                               int64_t ld, mclcar_samples** out) {
    if (!ctx || !M_dev) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    mclcar_samples* s = nullptr;
    MCL_TRY(make_samples(ctx, dtype, layout, sites, N, ld, &s));
    s->d_M = const_cast<void*>(M_dev);
    s->owned = false;
    *out = s;
    return MCLCAR_OK;
}

void mclcar_samples_free(mclcar_samples* s) {
    if (!s) return;
    mclcar_ctx* ctx = s->ctx;
    if (ctx && !ctx->host_only) {
        cudaSetDevice(ctx->device);
        if (s->owned && s->d_M) pool_free(ctx, s->d_M);
        if (s->d_q) pool_free(ctx, s->d_q);
        if (s->d_base) pool_free(ctx, s->d_base);
    }
    delete s;
}

int64_t mclcar_samples_count(const mclcar_samples* s) { return s ? s->N : 0; }

int mclcar_samples_fetch(mclcar_ctx* ctx, mclcar_samples* s, int layout, double* out, int64_t ld) {
    if (!ctx || !s || !out) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    if (!s->d_M) return fail(ctx, MCLCAR_ESTATE, "this sample set keeps sufficient statistics only (keep = 0)");
    const int64_t rows_src = s->layout == MCLCAR_COL_IS_SAMPLE ? s->N : s->sites;
    const int64_t cols_src = s->layout == MCLCAR_COL_IS_SAMPLE ? s->sites : s->N;
    const bool same = layout == s->layout;
    const int64_t rows = same ? rows_src : cols_src, cols = same ? cols_src : rows_src;
    if (ld < cols) return fail(ctx, MCLCAR_EINVAL, "ld too small");
    double* tmp = nullptr;
    MCL_TRY(pool_alloc(ctx, (void**)&tmp, (size_t)rows * cols * sizeof(double)));
    cudaStream_t st = ctx->stream;
    int status = MCLCAR_OK;
    if (same) {
        dim3 grid((unsigned)((cols + 255) / 256), (unsigned)rows);
        if (rows > 65535) {
            // one launch per block of rows
            for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
                const int64_t nr = std::min<int64_t>(65535, rows - r0);
                dim3 g((unsigned)((cols + 255) / 256), (unsigned)nr);
                if (s->dtype == MCLCAR_F32)
                    to_f64_kernel<float><<<g, 256, 0, st>>>((const float*)s->d_M + r0 * s->ld, tmp + r0 * cols, nr, cols, s->ld, cols);
                else
                    to_f64_kernel<double><<<g, 256, 0, st>>>((const double*)s->d_M + r0 * s->ld, tmp + r0 * cols, nr, cols, s->ld, cols);
            }
        } else if (s->dtype == MCLCAR_F32) {
            to_f64_kernel<float><<<grid, 256, 0, st>>>((const float*)s->d_M, tmp, rows, cols, s->ld, cols);
        } else {
            to_f64_kernel<double><<<grid, 256, 0, st>>>((const double*)s->d_M, tmp, rows, cols, s->ld, cols);
        }
    } else {
        // transpose in the source dtype, then widen
        void* t2 = nullptr;
        const size_t es = s->dtype == MCLCAR_F32 ? 4 : 8;
        if (pool_alloc(ctx, &t2, (size_t)rows * cols * es) != MCLCAR_OK) {
            pool_free(ctx, tmp);
            return MCLCAR_ENOMEM;
        }
        status = launch_transpose(ctx, s->d_M, s->dtype, rows_src, cols_src, s->ld, t2, cols);
        if (status == MCLCAR_OK) {
            for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
                const int64_t nr = std::min<int64_t>(65535, rows - r0);
                dim3 g((unsigned)((cols + 255) / 256), (unsigned)nr);
                if (s->dtype == MCLCAR_F32)
                    to_f64_kernel<float><<<g, 256, 0, st>>>((const float*)t2 + r0 * cols, tmp + r0 * cols, nr, cols, cols, cols);
                else
                    to_f64_kernel<double><<<g, 256, 0, st>>>((const double*)t2 + r0 * cols, tmp + r0 * cols, nr, cols, cols, cols);
            }
        }
        pool_free(ctx, t2);
    }
    if (status == MCLCAR_OK) {
        cudaError_t e = cudaMemcpy2DAsync(out, (size_t)ld * 8, tmp, (size_t)cols * 8, (size_t)cols * 8, (size_t)rows,
                                          cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) status = fail(ctx, MCLCAR_ECUDA, "device to host copy failed: %s", cudaGetErrorString(e));
    }
    pool_free(ctx, tmp);
    return status;
}

int mclcar_samples_diag(mclcar_ctx* ctx, mclcar_samples* s, double* accept_rate, int64_t* sweeps) {
    if (!ctx || !s) return MCLCAR_EINVAL;
    if (accept_rate) *accept_rate = s->accept;
    if (sweeps) *sweeps = s->sweeps;
    return MCLCAR_OK;
}

/* Effective sample size of each row of q ([d][ldq], sample index = e * chains + chain as both samplers
 * write them): pooled autocovariances along the chains, Geyer's initial positive sequence
 * (rho_0 + rho_1) + (rho_2 + rho_3) + ... summed while the pairs stay positive, ESS = N / tau with
 * tau = -1 + 2 * sum.  Pure host code: the reference draws independent samples (ESS = N), so this is
 * the number to look at before trusting `v.lr` from a Gibbs run. */
int mclcar_ess_from_stats(const double* q, int64_t N, int d, int64_t ldq, int64_t chains, double* ess_out) {
    if (!q || !ess_out || N < 1 || d < 1 || ldq < N || chains < 1) return MCLCAR_EINVAL;
    const int64_t C = chains < N ? chains : N;
    const int64_t E = (N + C - 1) / C;                      // states per chain (the last row of states may be partial)
    const int64_t maxlag = E - 1 < 200 ? E - 1 : 200;
    for (int j = 0; j < d; ++j) {
        const double* x = q + (size_t)j * ldq;
        long double m = 0;
        for (int64_t i = 0; i < N; ++i) m += x[i];
        m /= N;
        auto gamma = [&](int64_t k) {
            long double a = 0;
            int64_t cnt = 0;
            for (int64_t i = 0; i + k * C < N; ++i) {       // same chain, k states later: index + k * C
                a += ((long double)x[i] - m) * ((long double)x[i + k * C] - m);
                ++cnt;
            }
            return cnt ? (double)(a / cnt) : 0.0;
        };
        const double g0 = gamma(0);
        if (!(g0 > 0)) { ess_out[j] = (double)N; continue; }
        double sum = 0.0;
        for (int64_t k = 0; k <= maxlag; k += 2) {
            const double pair = gamma(k) / g0 + (k + 1 <= maxlag ? gamma(k + 1) / g0 : 0.0);
            if (pair <= 0) break;
            sum += pair;
        }
        double tau = -1.0 + 2.0 * sum;
        if (tau < 1e-3) tau = 1e-3;
        ess_out[j] = (double)N / tau;
    }
    return MCLCAR_OK;
}

/* the same for the statistics of a sample set drawn by one of the samplers (x'x, x'Wx, ... rows of q) */
int mclcar_samples_ess(mclcar_ctx* ctx, mclcar_samples* s, double* ess_out) {
    if (!ctx || !s || !ess_out) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    if (!s->stats_ready || !s->d_q || s->d < 1) return fail(ctx, MCLCAR_ESTATE, "statistics not computed yet: evaluate once (or call mclcar_dcar_stats) first");
    if (s->chains < 1) return fail(ctx, MCLCAR_ESTATE, "the sample matrix came from the caller: no chain structure");
    std::vector<double> q((size_t)s->d * s->N);
    MCL_CUDA(ctx, cudaMemcpyAsync(q.data(), s->d_q, q.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return mclcar_ess_from_stats(q.data(), s->N, s->d, s->N, s->chains, ess_out);
}

/* Retained sufficient statistics.  The reference keeps every iteration's sample matrix in its
 * result object (`MC.datas`, `Sim.data`: R/OptimMCL.R:80-93, R/rsmMCL.R:252-270) only so that summary() can call
 * vmle.dCAR / mc.gradlik.dCAR on it again (R/summary.OptimMCL.R:24-26, R/summary.rsmMCL.R:20-22).  Every estimator of
 * the path is a function of the N x d statistics alone, so a result object keeps those (N * d doubles instead of
 * N * n) and re-creates a sample set from them -- in a later R session, or on another GPU.
 * stats_export: q_out N x d column-major (as mclcar_dcar_stats); *d_out = d; q_out NULL just asks for d.  The
 * statistics must have been formed (any evaluation, mclcar_dcar_stats or mclcar_hcar_stats does it). */
int mclcar_samples_stats_export(mclcar_ctx* ctx, mclcar_samples* s, int* d_out, double* q_out) {
    if (!ctx || !s) return MCLCAR_EINVAL;
    if (s->ctx != ctx) return fail(ctx, MCLCAR_EINVAL, "sample set belongs to another context");
    MCL_TRY(require_device(ctx));
    if (!s->stats_ready || !s->d_q || s->d < 1)
        return fail(ctx, MCLCAR_ESTATE, "statistics not computed yet: evaluate once (or call mclcar_dcar_stats / mclcar_hcar_stats) first");
    if (d_out) *d_out = s->d;
    if (q_out) {
        MCL_CUDA(ctx, cudaMemcpyAsync(q_out, s->d_q, (size_t)s->d * s->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MCLCAR_OK;
}

/* a statistics-only sample set (no matrix: fetch fails, every likelihood entry works) from exported statistics;
 * `sites` = n (direct CAR) or K (hierarchical model), d = 2 + 2p or 4 + 2p; `chains` = the sampler's parallel chains
 * (mclcar_samples_schedule) or 0 for independent draws.  Attach the importance point afterwards as usual. */
int mclcar_samples_from_stats(mclcar_ctx* ctx, const double* q, int64_t N, int d, int64_t sites, int64_t chains,
                              mclcar_samples** out) {
    if (!ctx || !q || !out) return MCLCAR_EINVAL;
    *out = nullptr;
    MCL_TRY(require_device(ctx));
    if (d < 2 || d > 4 + 2 * MCLCAR_MAX_COV) return fail(ctx, MCLCAR_EINVAL, "bad number of statistics %d", d);
    for (int64_t i = 0; i < (int64_t)d * N; ++i)
        if (!std::isfinite(q[i])) return fail(ctx, MCLCAR_EINVAL, "non-finite statistic at position %lld", (long long)i);
    mclcar_samples* s = nullptr;
    MCL_TRY(make_samples(ctx, MCLCAR_F64, MCLCAR_COL_IS_SAMPLE, sites, N, sites, &s));
    int st = samples_alloc_stats(ctx, s, d);
    if (st == MCLCAR_OK) {
        cudaError_t e = cudaMemcpyAsync(s->d_q, q, (size_t)d * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = fail(ctx, MCLCAR_ECUDA, "upload of the statistics failed: %s", cudaGetErrorString(e));
    }
    if (st == MCLCAR_OK) st = samples_update_qbar(ctx, s);
    if (st != MCLCAR_OK) {
        mclcar_samples_free(s);
        return st;
    }
    s->stats_ready = true;
    s->chains = chains > 0 ? chains : 0;
    *out = s;
    return MCLCAR_OK;
}

int mclcar_samples_schedule(mclcar_ctx* ctx, mclcar_samples* s, int64_t* chains, int* burn, int* thin, double* omega) {
    if (!ctx || !s) return MCLCAR_EINVAL;
    if (chains) *chains = s->chains;
    if (burn) *burn = s->burn;
    if (thin) *thin = s->thin;
    if (omega) *omega = s->omega;
    return MCLCAR_OK;
}

}  // extern "C"

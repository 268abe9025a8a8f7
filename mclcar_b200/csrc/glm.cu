// Binomial / Poisson GLM with latent CAR effects: mirrors R/mcl.glm.R, R/mcl.bin.R and
// R/mcl.pois.R of the reference.
//
// For a latent sample z (row of Zy) and a candidate theta = (rho, sigma, beta):
//   logH(z; theta) = Const(theta) + cterm + D(z; beta) - (z'z - rho z'Wz)/(2 sigma)      (R/mcl.bin.R:54-59)
//   Const = -n/2 log(2 pi) + 1/2 logdet((I - rho W)/sigma),  cterm = sum lchoose | -sum lfactorial,
//   D     = y'eta - sum n_j log(1 + e^eta_j)   |   y'eta - sum e^eta_j,          eta = X beta + z
//   grad.logH = ( z'Wz/(2 sigma) - 1/2 sum lam/(1 - rho lam),  -n/(2 sigma) + z'(I - rho W)z/(2 sigma^2),
//                 X'(y - mu) )                                                         (R/mcl.bin.R:188-196)
//   Hessian.logH: (R/mcl.bin.R:260-274; Poisson beta-beta block as coded at R/mcl.pois.R:268)
// so every per-sample quantity is affine in the statistics
//   q = ( z'z, z'Wz,  { D_b, g_b[p], H_b[p(p+1)/2] : b = distinct beta of the batch } ).
// z'z and z'Wz come from the site-major statistics kernel (stats_csr.cu); the beta-dependent
// data terms from glm_terms_kernel below: ONE pass over the sample matrix per tile of
// candidate betas, lane = sample, fp64 exp / log1p.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "engine.hpp"

namespace mclcar {
namespace {

constexpr int kThreads = 128;
// candidate betas per pass over the sample matrix: all betas of a tile share the load / conversion of z and of the
// site's y, n, X row.  The tile is as wide as the accumulators allow: 8 for log-ratio batches (one accumulator per
// beta), 4 with gradients, 2 with Hessian blocks
constexpr int tile_of(int mode) { return mode == 0 ? 8 : (mode == 1 ? 4 : 2); }

// ---- binomial data term without exp / log1p / divide ---------------------------------------------------------
// log(1 + e^eta) = max(eta, 0) + L(|eta|),  L(a) = log1p(e^-a):  L is analytic with its nearest singularities at
// a = +- i pi, so a degree-8 polynomial per interval of width 1/4 (Chebyshev interpolation, Bernstein-ellipse
// parameter ~50) is exact to 1.2e-15 absolute on [0, 40) (L(40) = 4e-18); its derivative gives
// expit(a) = 1 + L'(a) to 5e-13.  One table row (9 doubles, shared memory) and 8 (16 with the derivative) DFMA
// replace the ~65 fp64 instructions of exp + log1p + divide: the binomial likelihood pass is bound by exactly
// that arithmetic (scripts/ubench/fp64_peaks.cu).  Rows are 9 doubles = 18 banks apart: lanes whose |eta| differ
// by less than 4 read conflict-free, equal intervals broadcast.
constexpr int kSpDeg = 8, kSpRow = kSpDeg + 1, kSpIntervals = 160;      // [0, 40) in steps of 1/4
constexpr double kSpMax = 39.999999;

const std::vector<double>& softplus_table() {
    static const std::vector<double> tab = [] {
        std::vector<double> t((size_t)kSpIntervals * kSpRow);
        const int m = kSpRow;
        for (int k = 0; k < kSpIntervals; ++k) {
            const long double c = (k + 0.5L) * 0.25L;
            long double f[kSpRow], a[kSpRow];
            for (int j = 0; j < m; ++j) f[j] = log1pl(expl(-(c + 0.125L * cosl(M_PIl * (j + 0.5L) / m))));
            for (int d = 0; d < m; ++d) {
                long double acc = 0;
                for (int j = 0; j < m; ++j) acc += f[j] * cosl(M_PIl * d * (j + 0.5L) / m);
                a[d] = acc * (d == 0 ? 1.0L : 2.0L) / m;
            }
            // Chebyshev -> monomial in t in [-1, 1]: T_0 = 1, T_1 = t, T_{d+1} = 2 t T_d - T_{d-1}
            long double mono[kSpRow] = {0}, T0[kSpRow] = {0}, T1[kSpRow] = {0}, T2[kSpRow];
            T0[0] = 1; T1[1] = 1;
            for (int i = 0; i < m; ++i) mono[i] += a[0] * T0[i] + (m > 1 ? a[1] * T1[i] : 0);
            for (int d = 2; d < m; ++d) {
                for (int i = 0; i < m; ++i) T2[i] = (i > 0 ? 2 * T1[i - 1] : 0) - T0[i];
                for (int i = 0; i < m; ++i) { mono[i] += a[d] * T2[i]; T0[i] = T1[i]; T1[i] = T2[i]; }
            }
            for (int i = 0; i < m; ++i) t[(size_t)k * m + i] = (double)mono[i];
        }
        return t;
    }();
    return tab;
}

// L(|eta|) and (WANT_D) dL/da at a = |eta| from the table (host and device: the same arithmetic)
template <bool WANT_D>
__host__ __device__ __forceinline__ double softplus_tail(const double* __restrict__ tab, double eta, double& dL) {
    const double a = fmin(fabs(eta), kSpMax);
    const int k = (int)(a * 4.0);
    const double t = fma(a, 8.0, -(double)(2 * k + 1));
    const double* c = tab + k * kSpRow;
    double p = c[kSpDeg], dp = 0.0;
#pragma unroll
    for (int j = kSpDeg - 1; j >= 0; --j) {
        if (WANT_D) dp = fma(dp, t, p);
        p = fma(p, t, c[j]);
    }
    dL = 8.0 * dp;
    return p;
}

// The candidate betas of a tile are close (line-search points, Jacobian columns): for one (sample, site) their eta
// almost always fall into the same table interval.  All betas of the tile are therefore evaluated together: one row
// read serves them all and their Horner chains interleave (eight independent DFMA chains instead of one dependent
// chain per beta behind a branch); only the lanes where the intervals differ read a row per beta.
template <int BT, bool WANT_D>
__device__ __forceinline__ void softplus_tail_tile(const double* __restrict__ tab, const double (&eta)[BT], double (&L)[BT],
                                                   double (&dL)[BT]) {
    double t[BT];
    int k[BT];
    bool same = true;
#pragma unroll
    for (int b = 0; b < BT; ++b) {
        double a = fabs(eta[b]);
        if (__double2hiint(a) >= 0x40440000) a = kSpMax;        // a >= 40 (or NaN): L = 0 to 4e-18; integer compare of the high word
        k[b] = (int)(a * 4.0);
        t[b] = fma(a, 8.0, -(double)(2 * k[b] + 1));
        same = same && k[b] == k[0];
    }
    double p[BT], dp[BT];
    if (same) {
        double c[kSpRow];
#pragma unroll
        for (int j = 0; j < kSpRow; ++j) c[j] = tab[k[0] * kSpRow + j];
#pragma unroll
        for (int b = 0; b < BT; ++b) { p[b] = c[kSpDeg]; dp[b] = 0.0; }
#pragma unroll
        for (int j = kSpDeg - 1; j >= 0; --j)
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                if (WANT_D) dp[b] = fma(dp[b], t[b], p[b]);
                p[b] = fma(p[b], t[b], c[j]);
            }
    } else {
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            const double* c = tab + k[b] * kSpRow;
            p[b] = c[kSpDeg]; dp[b] = 0.0;
#pragma unroll
            for (int j = kSpDeg - 1; j >= 0; --j) {
                if (WANT_D) dp[b] = fma(dp[b], t[b], p[b]);
                p[b] = fma(p[b], t[b], c[j]);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < BT; ++b) { L[b] = p[b]; dL[b] = 8.0 * dp[b]; }
}

struct GlmTermsParams {
    const double* sptab;  // softplus table (device), kSpIntervals x kSpRow
    const void* M;  // [n][ld] site-major
    int64_t ld, N;
    int n, p;
    const double* X;   // [p][n]
    const double* y;   // [n]
    const double* nt;  // [n] or null (== 1)
    const double* beta;  // [nb][p] this tile's betas
    int nb;              // betas in this tile (<= tile_of(mode))
    int chunk;           // sites per chunk
    double* out;         // [chunks][nb * ns][ldq]
    int64_t ldq;
    int ns;              // statistics per beta: 1 (+ p (+ p(p+1)/2))
    // STATS: the same pass also forms z'z and z'Wz (rows of `sout`, [chunks][2][ldq]) -- the whole GLM likelihood
    // pass reads the sample matrix once (upper-triangle edge list of W as in stats_csr.cu)
    const int* eptr;
    const int* ei;
    const int* ej;
    const double* ew;
    double* sout;
};

// MODE 0: D only; 1: D + gradient; 2: D + gradient + Hessian block
template <typename T, int FAM, int PT, int MODE, bool STATS, int BT>
__global__ void __launch_bounds__(kThreads) glm_terms_kernel(const GlmTermsParams prm) {
    constexpr int NH = PT * (PT + 1) / 2;
    constexpr int NS = 1 + (MODE >= 1 ? PT : 0) + (MODE >= 2 ? NH : 0);
    __shared__ double sbeta[BT * PT];
    __shared__ double stab[FAM == FAM_BINOM ? kSpIntervals * kSpRow : 1];
    if (threadIdx.x < BT * PT) {
        const int b = threadIdx.x / PT, a = threadIdx.x % PT;
        sbeta[threadIdx.x] = (b < prm.nb && a < prm.p) ? prm.beta[b * prm.p + a] : 0.0;
    }
    if (FAM == FAM_BINOM)
        for (int t = threadIdx.x; t < kSpIntervals * kSpRow; t += blockDim.x) stab[t] = __ldg(prm.sptab + t);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.N) return;
    const T* M = reinterpret_cast<const T*>(prm.M);
    const int s0 = blockIdx.y * prm.chunk, s1 = min(prm.n, s0 + prm.chunk);
    double acc[BT][NS];
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int t = 0; t < NS; ++t) acc[b][t] = 0.0;
    double zz = 0.0;
    for (int s = s0; s < s1; ++s) {
        const double z = (double)ld_stream(M + (size_t)s * prm.ld + i);
        const double ys = __ldg(prm.y + s);
        const double ns = prm.nt ? __ldg(prm.nt + s) : 1.0;
        if (STATS) zz = fma(z, z, zz);
        double xr[PT];
#pragma unroll
        for (int a = 0; a < PT; ++a) xr[a] = a < prm.p ? __ldg(prm.X + (size_t)a * prm.n + s) : 0.0;
        double etas[BT], Lt[BT], dLt[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) {          // slots beyond nb hold beta = 0: computed, never stored
            etas[b] = z;
#pragma unroll
            for (int a = 0; a < PT; ++a) etas[b] = fma(xr[a], sbeta[b * PT + a], etas[b]);
        }
        if (FAM == FAM_BINOM) softplus_tail_tile<BT, (MODE >= 1)>(stab, etas, Lt, dLt);
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            if (b < prm.nb) {
                const double eta = etas[b];
                double mu, v;
                if (FAM == FAM_BINOM) {
                    // log(1 + e^eta) = max(eta, 0) + L(|eta|), max(eta, 0) = (eta + |eta|)/2;
                    // p = expit(eta) = 1 + L'(|eta|) for eta >= 0, -L'(|eta|) below
                    acc[b][0] += ys * eta - ns * (0.5 * (eta + fabs(eta)) + Lt[b]);
                    if (MODE >= 1) {
                        const double pz = eta >= 0.0 ? 1.0 + dLt[b] : -dLt[b];
                        mu = ns * pz;
                        v = -ns * pz * (1.0 - pz);   // -n p (1 - p): R/mcl.bin.R:262,268
                    }
                } else {
                    const double ex = exp(eta);
                    acc[b][0] += ys * eta - ex;
                    mu = ex;
                    v = ex * ex;                 // +exp(eta)^2, as coded at R/mcl.pois.R:268
                }
                if (MODE >= 1) {
                    const double r = ys - mu;
#pragma unroll
                    for (int a = 0; a < PT; ++a) acc[b][1 + a] = fma(xr[a], r, acc[b][1 + a]);
                }
                if (MODE >= 2) {
                    int t = 1 + PT;
#pragma unroll
                    for (int a = 0; a < PT; ++a) {
                        const double va = v * xr[a];
#pragma unroll
                        for (int c = a; c < PT; ++c) {
                            acc[b][t] = fma(va, xr[c], acc[b][t]);
                            ++t;
                        }
                    }
                }
            }
        }
    }
    if (STATS) {
        // z'Wz / 2 over the chunk's edges: independent iterations, many gathers in flight (the rows were just streamed:
        // they come from L2)
        double zwz = 0.0;
        const int e0 = prm.eptr[s0], e1 = prm.eptr[s1];
#pragma unroll 4
        for (int e = e0; e < e1; ++e) {
            const double a = (double)__ldg(M + (size_t)__ldg(prm.ei + e) * prm.ld + i);
            const double b = (double)__ldg(M + (size_t)__ldg(prm.ej + e) * prm.ld + i);
            zwz = fma(prm.ew ? __ldg(prm.ew + e) * a : a, b, zwz);
        }
        double* so = prm.sout + (size_t)blockIdx.y * 2 * prm.ldq + i;
        so[0] = zz;
        so[prm.ldq] = 2.0 * zwz;
    }
    // compact (exact p) output order: D, g[0..p), H packed upper over p
    const int p = prm.p;
    double* o = prm.out + (size_t)blockIdx.y * prm.nb * prm.ns * prm.ldq + i;
#pragma unroll
    for (int b = 0; b < BT; ++b) {
        if (b < prm.nb) {
            double* ob = o + (size_t)b * prm.ns * prm.ldq;
            ob[0] = acc[b][0];
            if (MODE >= 1) {
#pragma unroll
                for (int a = 0; a < PT; ++a)
                    if (a < p) ob[(size_t)(1 + a) * prm.ldq] = acc[b][1 + a];
            }
            if (MODE >= 2) {
                int t = 1 + PT, tc = 1 + p;
#pragma unroll
                for (int a = 0; a < PT; ++a)
#pragma unroll
                    for (int c = a; c < PT; ++c) {
                        if (a < p && c < p) ob[(size_t)(tc++) * prm.ldq] = acc[b][t];
                        ++t;
                    }
            }
        }
    }
}

__global__ void glm_chunk_sum_kernel(const double* partial, double* q, int S, int d, int64_t ldq, int64_t N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int j = 0; j < d; ++j) {
        double a = 0.0;
        for (int s = 0; s < S; ++s) a += partial[((size_t)s * d + j) * ldq + i];
        q[(size_t)j * ldq + i] = a;
    }
}

template <typename T, int FAM, int PT, bool STATS>
void launch_mode(const GlmTermsParams& p, int mode, dim3 grid, cudaStream_t st) {
    // a tile's slots are all evaluated: log-ratio passes with one or two betas (lHZy.psi, the tail of a batch) take the
    // two-slot build
    if (mode == 0 && p.nb <= 2) glm_terms_kernel<T, FAM, PT, 0, STATS, 2><<<grid, kThreads, 0, st>>>(p);
    else if (mode == 0) glm_terms_kernel<T, FAM, PT, 0, STATS, tile_of(0)><<<grid, kThreads, 0, st>>>(p);
    else if (mode == 1) glm_terms_kernel<T, FAM, PT, 1, STATS, tile_of(1)><<<grid, kThreads, 0, st>>>(p);
    else glm_terms_kernel<T, FAM, PT, 2, STATS, tile_of(2)><<<grid, kThreads, 0, st>>>(p);
}
template <typename T, int FAM, bool STATS>
void launch_p(const GlmTermsParams& p, int mode, dim3 grid, cudaStream_t st) {
    if (p.p <= 1) launch_mode<T, FAM, 1, STATS>(p, mode, grid, st);
    else if (p.p <= 2) launch_mode<T, FAM, 2, STATS>(p, mode, grid, st);
    else if (p.p <= 4) launch_mode<T, FAM, 4, STATS>(p, mode, grid, st);
    else launch_mode<T, FAM, 8, STATS>(p, mode, grid, st);
}
template <typename T, int FAM>
void launch_s(const GlmTermsParams& p, int mode, dim3 grid, cudaStream_t st) {
    if (p.sout) launch_p<T, FAM, true>(p, mode, grid, st);
    else launch_p<T, FAM, false>(p, mode, grid, st);
}

// data terms of `nb_total` distinct betas -> rows [row0 + b*ns, ...) of q [.][ldq]; with d_static != null the first
// pass also writes z'z, z'Wz there ([2][ldq]): ONE read of the sample matrix for the whole likelihood pass when the
// batch fits one beta tile
int glm_data_terms(mclcar_ctx* ctx, const mclcar_samples* s, int fam, const std::vector<double>& betas, int nb_total,
                   int mode, double* d_q, int64_t ldq, int row0, double* d_static) {
    const int p = ctx->p, n = ctx->n;
    const int ns = 1 + (mode >= 1 ? p : 0) + (mode >= 2 ? p * (p + 1) / 2 : 0);
    const int BT = tile_of(mode);
    const int64_t N = s->N;
    const int64_t xblocks = (N + kThreads - 1) / kThreads;
    // site chunks: ~16 CTAs per SM in the grid (six resident at 80 registers): the round-1 grid of 4.3 CTAs per SM left a
    // fifth of the machine idle in its last wave; measured on config 3 (79 sample blocks): 8 chunks 6.4 ms per step,
    // 15: 4.8, 30: 4.4, 45-60: 4.4
    int chunks = (int)std::max<int64_t>(1, ((int64_t)16 * ctx->n_sm + xblocks / 2) / xblocks);
    if (const char* e = getenv("MCLCAR_GLM_CHUNKS")) chunks = std::max(1, atoi(e));
    chunks = std::min(chunks, std::max(1, n / 32));
    const int chunk = (n + chunks - 1) / chunks;
    chunks = (n + chunk - 1) / chunk;
    const EdgeCache* ec = nullptr;
    if (d_static) MCL_TRY(edge_cache_get(ctx, n, ctx->rowptr, ctx->colidx, ctx->val, &ec));
    if (fam == FAM_BINOM && !ctx->d_sptab.p) {
        const std::vector<double>& tab = softplus_table();
        MCL_TRY(ensure(ctx, ctx->d_sptab, tab.size() * sizeof(double)));
        MCL_CUDA(ctx, cudaMemcpyAsync(ctx->d_sptab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    double* d_beta = nullptr;
    MCL_TRY(pool_alloc(ctx, (void**)&d_beta, betas.size() * sizeof(double)));
    MCL_CUDA(ctx, cudaMemcpyAsync(d_beta, betas.data(), betas.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double *partial = nullptr, *spartial = nullptr;
    if (chunks > 1) MCL_TRY(pool_alloc(ctx, (void**)&partial, (size_t)chunks * BT * ns * ldq * sizeof(double)));
    if (chunks > 1 && d_static) MCL_TRY(pool_alloc(ctx, (void**)&spartial, (size_t)chunks * 2 * ldq * sizeof(double)));
    int st = MCLCAR_OK;
    for (int b0 = 0; b0 < nb_total && st == MCLCAR_OK; b0 += BT) {
        GlmTermsParams prm{};
        prm.M = s->d_M; prm.ld = s->ld; prm.N = N; prm.n = n; prm.p = p;
        prm.sptab = (const double*)ctx->d_sptab.p;
        prm.X = ctx->d_X; prm.y = ctx->d_y; prm.nt = ctx->d_ntrial;
        prm.beta = d_beta + (size_t)b0 * p; prm.nb = std::min(BT, nb_total - b0);
        prm.chunk = chunk; prm.ldq = ldq; prm.ns = ns;
        double* dst = d_q + (size_t)(row0 + b0 * ns) * ldq;
        prm.out = chunks > 1 ? partial : dst;
        const bool with_stats = d_static && b0 == 0;
        if (with_stats) {
            prm.eptr = ec->d_eptr; prm.ei = ec->d_ei; prm.ej = ec->d_ej; prm.ew = ec->d_ew;
            prm.sout = chunks > 1 ? spartial : d_static;
        }
        dim3 grid((unsigned)xblocks, (unsigned)chunks);
        ProfScope prof(ctx, PROF_GLM, (chunks > 1 ? 2 : 1) + (with_stats && chunks > 1 ? 1 : 0));
        const bool f32 = s->dtype == MCLCAR_F32;
        if (fam == FAM_BINOM) {
            if (f32) launch_s<float, FAM_BINOM>(prm, mode, grid, ctx->stream);
            else launch_s<double, FAM_BINOM>(prm, mode, grid, ctx->stream);
        } else {
            if (f32) launch_s<float, FAM_POIS>(prm, mode, grid, ctx->stream);
            else launch_s<double, FAM_POIS>(prm, mode, grid, ctx->stream);
        }
        if (chunks > 1) {
            glm_chunk_sum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(partial, dst, chunks, prm.nb * ns, ldq, N);
            if (with_stats)
                glm_chunk_sum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(spartial, d_static, chunks, 2, ldq, N);
        }
        if (cudaGetLastError() != cudaSuccess) st = fail(ctx, MCLCAR_ECUDA, "GLM data-term kernel launch failed");
    }
    pool_free(ctx, partial);
    pool_free(ctx, spartial);
    pool_free(ctx, d_beta);
    return st;
}

// ---- host-side scalars -----------------------------------------------------------------
int need_eigs(mclcar_ctx* ctx) { return ensure_spectrum(ctx); }
double half_logdet_Q(const mclcar_ctx* ctx, double rho, double sigma) {
    long double a = 0;
    auto it = std::lower_bound(ctx->logdet_tab.begin(), ctx->logdet_tab.end(), std::make_pair(rho, (long double)-INFINITY));
    if (it != ctx->logdet_tab.end() && it->first == rho) a = it->second;
    else a = spectral_sum(ctx, [&](double l) { return (long double)std::log1p(-rho * l); });
    return (double)(0.5L * (a - (long double)ctx->n * std::log(sigma)));
}
double sum_lam_ratio(const mclcar_ctx* ctx, double rho, int power) {
    return (double)spectral_sum(ctx, [&](double l) {
        const double r = l / (1.0 - rho * l);
        return (long double)(power == 1 ? r : r * r);
    });
}
// sum log(1 - rho lam) of every distinct rho of a candidate batch in one go (exact.cu: one device launch when the sums
// are long) -- half_logdet_Q finds them in ctx->logdet_tab; with n = 5000 sites and 53 candidates the per-candidate
// host sums were 2.6 ms of a config-3 step.  The table is transient: the guard of the caller that filled it clears it.
struct LogdetTabGuard {
    mclcar_ctx* c;
    bool own;
    ~LogdetTabGuard() { if (own) c->logdet_tab.clear(); }
};
int fill_logdet_tab(mclcar_ctx* ctx, std::vector<double> rhos) {
    std::sort(rhos.begin(), rhos.end());
    rhos.erase(std::unique(rhos.begin(), rhos.end()), rhos.end());
    std::vector<double> ld(rhos.size());
    MCL_TRY(logdet_sums(ctx, rhos.data(), (int)rhos.size(), ld.data()));
    ctx->logdet_tab.clear();
    for (size_t i = 0; i < rhos.size(); ++i) ctx->logdet_tab.emplace_back(rhos[i], (long double)ld[i]);
    return MCLCAR_OK;
}
// psi-independent data constant: sum(sapply(y, lchoose, n = n.trial)) | -sum(lfactorial(y))
double glm_cterm(const mclcar_ctx* ctx, int fam) {
    const int n = ctx->n;
    long double a = 0;
    if (fam == FAM_POIS) {
        for (int j = 0; j < n; ++j) a -= std::lgamma(ctx->y[j] + 1.0);
        return (double)a;
    }
    auto lch = [](double nn, double k) { return std::lgamma(nn + 1.0) - std::lgamma(k + 1.0) - std::lgamma(nn - k + 1.0); };
    bool scalar = true;
    for (int j = 1; j < n && scalar; ++j) scalar = ctx->ntrial.empty() || ctx->ntrial[j] == ctx->ntrial[0];
    if (scalar) {
        const double nn = ctx->ntrial.empty() ? 1.0 : ctx->ntrial[0];
        for (int j = 0; j < n; ++j) a += lch(nn, ctx->y[j]);
    } else {  // vector n.trial: the reference's sapply sums the whole n x n table (R/mcl.bin.R:26)
        for (int j = 0; j < n; ++j)
            for (int l = 0; l < n; ++l) a += lch(ctx->ntrial[l], ctx->y[j]);
    }
    return (double)a;
}

struct GlmBatch {
    int fam = FAM_BINOM, p = 0, P = 0, K = 0;
    int nb = 0;                  // distinct betas (index 0 = the importance beta when the base is internal)
    std::vector<double> betas;   // nb x p
    std::vector<int> bidx;       // K: beta index of each candidate
    int mode = 0, ns = 1, d = 2; // statistics per beta, total rows
    double* d_q = nullptr;       // [d][N]: z'z, z'Wz, then per beta: D, g, H
    std::vector<double> qbar;
};

int glm_check(mclcar_ctx* ctx, mclcar_samples* s, int P) {
    if (!ctx || !s) return MCLCAR_EINVAL;
    if (s->ctx != ctx) return fail(ctx, MCLCAR_EINVAL, "sample set belongs to another context");
    if (ctx->kind == GRAPH_NONE || !ctx->has_obs) return fail(ctx, MCLCAR_ESTATE, "set graph, design and observations first");
    if (ctx->p < 1) return fail(ctx, MCLCAR_ESTATE, "the GLM needs a design matrix covX");
    if (P != 2 + ctx->p) return fail(ctx, MCLCAR_EINVAL, "parameter vector of length %d; expected %d", P, 2 + ctx->p);
    if (s->sites != ctx->n) return fail(ctx, MCLCAR_EINVAL, "sample matrix has %lld sites, graph has %d", (long long)s->sites, ctx->n);
    if (s->layout != MCLCAR_ROW_IS_SAMPLE) return fail(ctx, MCLCAR_EINVAL, "GLM sample sets are N x n with row = sample (Zy)");
    if (!s->d_M) return fail(ctx, MCLCAR_ESTATE, "GLM sample sets must keep the sample matrix");
    return need_eigs(ctx);
}

int glm_build_batch(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int mode, GlmBatch& gb) {
    const int p = ctx->p;
    gb.fam = s->family; gb.p = p; gb.P = P; gb.K = K; gb.mode = mode;
    gb.ns = 1 + (mode >= 1 ? p : 0) + (mode >= 2 ? p * (p + 1) / 2 : 0);
    gb.betas.clear(); gb.bidx.assign(K, 0); gb.nb = 0;
    auto find_or_add = [&](const double* b) {
        for (int t = 0; t < gb.nb; ++t)
            if (std::memcmp(gb.betas.data() + (size_t)t * p, b, p * sizeof(double)) == 0) return t;
        gb.betas.insert(gb.betas.end(), b, b + p);
        return gb.nb++;
    };
    if (!s->d_base) find_or_add(s->psi0.data() + 2);  // index 0: importance beta
    for (int k = 0; k < K; ++k) gb.bidx[k] = find_or_add(psi + (size_t)k * P + 2);
    gb.d = 2 + gb.nb * gb.ns;
    // z'z and z'Wz stay with the sample set; the first evaluation forms them in the same pass as the data terms
    const bool have_static = s->stats_ready && s->d == 2;
    if (!have_static) MCL_TRY(samples_alloc_stats(ctx, s, 2));
    MCL_TRY(pool_alloc(ctx, (void**)&gb.d_q, (size_t)gb.d * s->N * sizeof(double)));
    MCL_TRY(glm_data_terms(ctx, s, gb.fam, gb.betas, gb.nb, mode, gb.d_q, s->N, 2, have_static ? nullptr : s->d_q));
    s->stats_ready = true;
    MCL_CUDA(ctx, cudaMemcpyAsync(gb.d_q, s->d_q, (size_t)2 * s->N * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    // column means (shift and centres)
    mclcar_samples tmp;
    tmp.N = s->N; tmp.d = gb.d; tmp.d_q = gb.d_q;
    MCL_TRY(samples_update_qbar(ctx, &tmp));
    gb.qbar = tmp.qbar;
    tmp.d_q = nullptr;
    return MCLCAR_OK;
}

// affine maps of candidate k over the batch statistics
struct GlmAffine {
    double c0 = 0;
    std::vector<double> c;      // d
    std::vector<double> a0, A;  // P, P x d  (gradient of logH)
    double rho = 0, sigma = 1;
};

void glm_affine(const mclcar_ctx* ctx, const GlmBatch& gb, const double* pars, int b, double cterm, GlmAffine& m) {
    const int d = gb.d, P = gb.P, p = gb.p, n = ctx->n;
    const double rho = pars[0], sigma = pars[1];
    m.rho = rho; m.sigma = sigma;
    m.c.assign(d, 0.0);
    m.c0 = -0.5 * n * std::log(2.0 * M_PI) + half_logdet_Q(ctx, rho, sigma) + cterm;
    m.c[0] = -1.0 / (2 * sigma);
    m.c[1] = rho / (2 * sigma);
    const int rb = 2 + b * gb.ns;  // first row of this beta's statistics
    m.c[rb] = 1.0;
    m.a0.assign(P, 0.0);
    m.A.assign((size_t)P * d, 0.0);
    if (gb.mode >= 1) {
        m.a0[0] = -0.5 * sum_lam_ratio(ctx, rho, 1);
        m.A[1] = 0.5 / sigma;
        m.a0[1] = -n / (2 * sigma);
        m.A[(size_t)d + 0] = 1.0 / (2 * sigma * sigma);
        m.A[(size_t)d + 1] = -rho / (2 * sigma * sigma);
        for (int a = 0; a < p; ++a) m.A[(size_t)(2 + a) * d + rb + 1 + a] = 1.0;
    }
}

int glm_partial(mclcar_ctx* ctx, mclcar_samples* s, GlmBatch& gb, const double* psi, int what, int shift_mode,
                const int64_t* subset_idx, const int64_t* subset_ptr, double* tuples) {
    const int d = gb.d, K = gb.K, P = gb.P;
    const int F = what == WHAT_LR ? 0 : P;
    ReducePlan plan;
    plan.d_q = gb.d_q; plan.ldq = s->N; plan.d = d; plan.N = s->N; plan.K = K; plan.F = F; plan.what = what;
    plan.shift_mode = shift_mode; plan.d_base = s->d_base;
    plan.subset_idx = subset_idx; plan.subset_ptr = subset_ptr;
    const int cs = coef_stride_of(d, F);
    plan.coef_stride = cs;
    plan.coef.assign((size_t)K * cs, 0.0);
    const double cterm = s->d_base ? glm_cterm(ctx, gb.fam) : 0.0;
    GlmAffine m0, m;
    if (!s->d_base) glm_affine(ctx, gb, s->psi0.data(), 0, 0.0, m0);
    for (int k = 0; k < K; ++k) {
        glm_affine(ctx, gb, psi + (size_t)k * P, gb.bidx[k], cterm, m);
        double* cf = plan.coef.data() + (size_t)k * cs;
        cf[0] = m.c0;
        for (int j = 0; j < d; ++j) cf[1 + j] = m.c[j];
        if (!s->d_base) {  // r = logH(z; theta) - logH(z; psi0), coefficients differenced first
            cf[0] -= m0.c0;
            for (int j = 0; j < d; ++j) cf[1 + j] -= m0.c[j];
        }
        double shift = cf[0];
        for (int j = 0; j < d; ++j) shift += cf[1 + j] * gb.qbar[j];
        if (s->d_base) shift -= s->base_mean;
        cf[coef_off_shift(d)] = shift;
        int64_t nsub = s->N;
        if (subset_ptr && subset_ptr[k + 1] > subset_ptr[k]) nsub = subset_ptr[k + 1] - subset_ptr[k];
        cf[coef_off_n(d)] = (double)nsub;
        cf[coef_off_centre(d)] = 1.0;
        for (int a = 0; a < F; ++a) {
            double* fa = cf + coef_off_feat(d, F) + (size_t)a * (1 + d);
            fa[0] = m.a0[a];
            double fbar = m.a0[a];
            for (int j = 0; j < d; ++j) {
                fa[1 + j] = m.A[(size_t)a * d + j];
                fbar += fa[1 + j] * gb.qbar[j];
            }
            cf[coef_off_centre(d) + 1 + a] = fbar;
        }
    }
    MCL_TRY(run_reduce(ctx, plan, tuples, true));
    return gather_and_merge(ctx, what, F, d, K, tuples);
}

}  // namespace
}  // namespace mclcar

using namespace mclcar;

extern "C" {

int mclcar_glm_attach(mclcar_ctx* ctx, mclcar_samples* s, int family, const double* psi0, int P, const double* lHZy_psi) {
    if (!psi0) return MCLCAR_EINVAL;
    if (family != MCLCAR_FAMILY_BINOM && family != MCLCAR_FAMILY_POISSON) return fail(ctx, MCLCAR_EINVAL, "family must be binomial or Poisson");
    MCL_TRY(glm_check(ctx, s, P));
    MCL_TRY(require_device(ctx));
    if (!(psi0[1] > 0)) return fail(ctx, MCLCAR_EINVAL, "sigma must be positive");
    s->family = family == MCLCAR_FAMILY_BINOM ? FAM_BINOM : FAM_POIS;
    s->psi0.assign(psi0, psi0 + P);
    if (s->d_base) { pool_free(ctx, s->d_base); s->d_base = nullptr; }
    s->base_mean = 0.0;
    if (lHZy_psi) {
        MCL_TRY(pool_alloc(ctx, (void**)&s->d_base, (size_t)s->N * sizeof(double)));
        MCL_CUDA(ctx, cudaMemcpyAsync(s->d_base, lHZy_psi, (size_t)s->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        long double a = 0;
        for (int64_t i = 0; i < s->N; ++i) a += lHZy_psi[i];
        s->base_mean = (double)(a / s->N);
    }
    s->attached = true;
    return MCLCAR_OK;
}

/* lHZy.psi_i = log f_psi0(y, Z_i), R/mcl.glm.R:34-37 */
int mclcar_glm_lhzy(mclcar_ctx* ctx, mclcar_samples* s, double* out) {
    if (!out || !s || !s->attached) return MCLCAR_EINVAL;
    MCL_TRY(glm_check(ctx, s, (int)s->psi0.size()));
    MCL_TRY(require_device(ctx));
    if (s->d_base) {
        MCL_CUDA(ctx, cudaMemcpyAsync(out, s->d_base, (size_t)s->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return MCLCAR_OK;
    }
    // the batch holds one "candidate" = psi0 itself: its beta is the importance beta (index 0)
    GlmBatch gb;
    const int P = (int)s->psi0.size();
    std::vector<double> psi(s->psi0);
    int st = glm_build_batch(ctx, s, psi.data(), 1, P, 0, gb);
    if (st != MCLCAR_OK) { pool_free(ctx, gb.d_q); return st; }
    std::vector<double> q((size_t)gb.d * s->N);
    cudaError_t e = cudaMemcpyAsync(q.data(), gb.d_q, q.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    pool_free(ctx, gb.d_q);
    if (e != cudaSuccess) return fail(ctx, MCLCAR_ECUDA, "copy of the statistics failed: %s", cudaGetErrorString(e));
    GlmAffine m;
    glm_affine(ctx, gb, psi.data(), gb.bidx[0], glm_cterm(ctx, gb.fam), m);
    for (int64_t i = 0; i < s->N; ++i) {
        double h = m.c0;
        for (int j = 0; j < gb.d; ++j) h += m.c[j] * q[(size_t)j * s->N + i];
        out[i] = h;
    }
    return MCLCAR_OK;
}

static int glm_run(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int what, int evar,
                   const int64_t* subset_idx, const int64_t* subset_ptr, int shift_mode, double* out0, double* out1) {
    if (!s || !psi || K < 1 || !out0) return MCLCAR_EINVAL;
    MCL_TRY(glm_check(ctx, s, P));
    MCL_TRY(require_device(ctx));
    if (!s->attached) return fail(ctx, MCLCAR_ESTATE, "call mclcar_glm_attach (or _prep) first");
    if (P > MCLCAR_MAX_P) return fail(ctx, MCLCAR_ELIMIT, "P = %d exceeds MCLCAR_MAX_P", P);
    for (int k = 0; k < K; ++k)
        if (!(psi[(size_t)k * P + 1] > 0)) return fail(ctx, MCLCAR_EINVAL, "sigma must be positive (candidate %d)", k);
    LogdetTabGuard tab_guard{ctx, ctx->logdet_tab.empty()};
    if (tab_guard.own) {
        std::vector<double> rhos(K + 1);
        rhos[0] = s->psi0[0];
        for (int k = 0; k < K; ++k) rhos[1 + k] = psi[(size_t)k * P];
        MCL_TRY(fill_logdet_tab(ctx, std::move(rhos)));
    }
    GlmBatch gb;
    const int mode = what == WHAT_LR ? 0 : (what == WHAT_GRAD ? 1 : 2);
    int st = glm_build_batch(ctx, s, psi, K, P, mode, gb);
    const int d = gb.d, F = what == WHAT_LR ? 0 : P, p = ctx->p, n = ctx->n;
    const int64_t T = tuple_len(what, F, d);
    std::vector<double> tuples((size_t)K * (T > 0 ? T : 1));
    if (st == MCLCAR_OK) st = glm_partial(ctx, s, gb, psi, what, shift_mode, subset_idx, subset_ptr, tuples.data());
    pool_free(ctx, gb.d_q);
    MCL_TRY(st);
    std::vector<double> vbar(1 + F), C((size_t)(1 + F) * (1 + F));
    for (int k = 0; k < K; ++k) {
        const double* tup = tuples.data() + (size_t)k * T;
        const double* pars = psi + (size_t)k * P;
        const double nn = tup[0], shift = tup[1];
        if (what == WHAT_LR) {  // R/mcl.bin.R:61-78
            centred_moments(tup, 1, vbar.data(), C.data());
            out0[k] = std::log(vbar[0]) + shift;
            if (out1) out1[k] = C[0] / (nn - 1.0) / (vbar[0] * vbar[0]) / nn;
        } else if (what == WHAT_GRAD) {  // R/mcl.bin.R:198-216
            const int V = 1 + P;
            centred_moments(tup, V, vbar.data(), C.data());
            const double eb = vbar[0];
            for (int a = 0; a < P; ++a) out0[(size_t)k * P + a] = vbar[1 + a] / eb;
            if (out1 && evar)
                for (int a = 0; a < P; ++a)
                    for (int b = 0; b < P; ++b)
                        out1[(size_t)k * P * P + a * P + b] = C[(size_t)(1 + a) * V + 1 + b] / (nn - 1.0) / (eb * eb) / nn;
        } else {  // R/mcl.bin.R:276-296
            const double rho = pars[0], sigma = pars[1];
            const double Se = tup[2];
            const double* Sf = tup + 3;
            const double* Sff = tup + 3 + P;
            const double* R1 = tup + 3 + P + P * (P + 1) / 2;
            const int rb = 2 + gb.bidx[k] * gb.ns;
            const double zz = R1[0] / Se, zWz = R1[1] / Se;  // weighted means of the statistics
            double* o = out0 + (size_t)k * P * P;
            for (int t = 0; t < P * P; ++t) o[t] = 0.0;
            // mean(w Hessian.logH)
            o[0] = -0.5 * sum_lam_ratio(ctx, rho, 2);
            o[1] = o[P] = -0.5 * zWz / (sigma * sigma);
            o[P + 1] = -((zz - rho * zWz) / sigma) / (sigma * sigma) + 0.5 * n / (sigma * sigma);
            int t = rb + 1 + p;
            for (int a = 0; a < p; ++a)
                for (int b = a; b < p; ++b, ++t) o[(2 + a) * P + 2 + b] = o[(2 + b) * P + 2 + a] = R1[t] / Se;
            // + mean(w g g') - (mean w g)(mean w g)'
            int sidx = 0;
            for (int a = 0; a < P; ++a)
                for (int b = a; b < P; ++b, ++sidx) {
                    const double v = Sff[sidx] / Se - (Sf[a] / Se) * (Sf[b] / Se);
                    o[a * P + b] += v;
                    if (a != b) o[b * P + a] += v;
                }
        }
    }
    return MCLCAR_OK;
}

int mclcar_glm_eval(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, const int64_t* subset_idx,
                    const int64_t* subset_ptr, int shift_mode, double* lr, double* var) {
    return glm_run(ctx, s, psi, K, P, WHAT_LR, 0, subset_idx, subset_ptr, shift_mode, lr, var);
}
int mclcar_glm_grad(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int evar, int shift_mode,
                    double* grad, double* gradvar) {
    return glm_run(ctx, s, psi, K, P, WHAT_GRAD, evar, nullptr, nullptr, shift_mode, grad, gradvar);
}
int mclcar_glm_hess(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int shift_mode, double* hess) {
    return glm_run(ctx, s, psi, K, P, WHAT_HESS, 0, nullptr, nullptr, shift_mode, hess, nullptr);
}

/* diagnostic: L(|eta|) = log1p(exp(-|eta|)) and its derivative from the table the binomial data-term kernel uses,
 * evaluated on the host with the kernel's arithmetic (pure host code; tests/test_abi_host.py checks it against libm) */
int mclcar_glm_softplus_table_eval(const double* eta, int64_t n, double* L_out, double* dL_out) {
    if (!eta || !L_out || n < 0) return MCLCAR_EINVAL;
    const std::vector<double>& tab = softplus_table();
    for (int64_t i = 0; i < n; ++i) {
        double d;
        L_out[i] = softplus_tail<true>(tab.data(), eta[i], d);
        if (dL_out) dL_out[i] = d;
    }
    return MCLCAR_OK;
}

/* Monte Carlo likelihood surface of the GLM: mcl.glm(c(rho_i, sigma_j, beta), mcdata, family, Evar = TRUE) on the grid
 * expand.grid(r.vec, s.vec), rho fastest (lr[i + nr * j]) -- the backdrop plot.rsmMCL has no exact values for when the
 * family is not Gaussian (R/rsmMCL.R:95-108 stops with "No exact value"; R/plot.rsmMCL.R:3).  At beta = the importance
 * beta (every rsm design point, R/rsmSub.R:169-174) the data terms cancel and the whole grid is ONE reduction over
 * (z'z, z'Wz); the log-determinants are formed once per distinct rho. */
int mclcar_glm_surface(mclcar_ctx* ctx, mclcar_samples* s, const double* rho, int nr, const double* sigma, int ns,
                       const double* beta, int shift_mode, double* lr, double* var) {
    if (!ctx || !s || !rho || !sigma || !beta || !lr || nr < 1 || ns < 1) return MCLCAR_EINVAL;
    const int p = ctx->p, P = 2 + p;
    MCL_TRY(glm_check(ctx, s, P));
    MCL_TRY(need_eigs(ctx));
    LogdetTabGuard guard{ctx, true};
    {
        std::vector<double> rhos(rho, rho + nr);
        rhos.push_back(s->psi0.empty() ? rho[0] : s->psi0[0]);
        MCL_TRY(fill_logdet_tab(ctx, std::move(rhos)));
    }
    const int64_t K = (int64_t)nr * ns, chunk = 16384;
    std::vector<double> psi((size_t)std::min(K, chunk) * P);
    for (int64_t k0 = 0; k0 < K; k0 += chunk) {
        const int kc = (int)std::min(chunk, K - k0);
        for (int k = 0; k < kc; ++k) {
            double* t = psi.data() + (size_t)k * P;
            t[0] = rho[(k0 + k) % nr];
            t[1] = sigma[(k0 + k) / nr];
            for (int a = 0; a < p; ++a) t[2 + a] = beta[a];
        }
        MCL_TRY(glm_run(ctx, s, psi.data(), kc, P, WHAT_LR, 0, nullptr, nullptr, shift_mode, lr + k0, var ? var + k0 : nullptr));
    }
    return MCLCAR_OK;
}

/* get.beta.glm (R/mcl.bin.R:83-95, R/mcl.glm.R:50-54) and the beta solve inside mcl.*.profile
 * (R/mcl.bin.R:111-123): root of the beta block of mcl.grad.glm at fixed (rho, sigma), with the
 * reference's penalty (sign(g) 1e5 once any |beta - beta0| > 2).  The reference hands the function
 * to nleqslv (CRAN, version unpinned, not vendored) with ftol = 1 and maxit = 2 or 3.  Restated
 * here: Newton iterations on a forward-difference Jacobian (step sqrt(eps) max(|beta_j|, 1)), stop as
 * soon as max|g| <= ftol, step halving (<= 4) when |g|^2 does not decrease.  The p + 1 gradient
 * evaluations of one Jacobian go through ONE batched mclcar_glm_grad call: the candidate betas
 * share the passes over the sample matrix. */
int mclcar_glm_get_beta(mclcar_ctx* ctx, mclcar_samples* s, const double* rho_sig, const double* beta0, int maxit,
                        double ftol, double* beta_out, double* fvec_out, int* iters_out, int* termcd_out) {
    if (!ctx || !s || !rho_sig || !beta0 || !beta_out) return MCLCAR_EINVAL;
    const int p = ctx->p, P = 2 + p;
    MCL_TRY(glm_check(ctx, s, P));
    if (maxit < 0) return fail(ctx, MCLCAR_EINVAL, "maxit must be >= 0");
    std::vector<double> x(beta0, beta0 + p), f(p), psi((size_t)(p + 1) * P), g((size_t)(p + 1) * P), J((size_t)p * p), dx(p);
    // the reference's grad.beta closure for a batch of betas (rows of B): F rows of length p
    auto F = [&](const double* B, int nb, double* out) -> int {
        for (int k = 0; k < nb; ++k) {
            psi[(size_t)k * P] = rho_sig[0];
            psi[(size_t)k * P + 1] = rho_sig[1];
            for (int a = 0; a < p; ++a) psi[(size_t)k * P + 2 + a] = B[(size_t)k * p + a];
        }
        MCL_TRY(glm_run(ctx, s, psi.data(), nb, P, WHAT_GRAD, 0, nullptr, nullptr, MCLCAR_SHIFT_MEAN, g.data(), nullptr));
        for (int k = 0; k < nb; ++k) {
            bool far = false;
            for (int a = 0; a < p; ++a) far = far || std::fabs(beta0[a] - B[(size_t)k * p + a]) > 2.0;
            for (int a = 0; a < p; ++a) {
                const double v = g[(size_t)k * P + 2 + a];
                out[(size_t)k * p + a] = far ? ((v > 0) - (v < 0)) * 1e5 : v;
            }
        }
        return MCLCAR_OK;
    };
    auto maxabs = [&](const std::vector<double>& v) { double m = 0; for (double t : v) m = std::max(m, std::fabs(t)); return m; };
    auto sumsq = [&](const std::vector<double>& v) { double m = 0; for (double t : v) m += t * t; return m; };
    MCL_TRY(F(x.data(), 1, f.data()));
    int it = 0, termcd = 4;   // nleqslv codes: 1 = function criterion met, 4 = iteration limit, 3 = no better point
    std::vector<double> B((size_t)(p + 1) * p), Fb((size_t)(p + 1) * p), xn(p), fn(p);
    while (true) {
        if (maxabs(f) <= ftol) { termcd = 1; break; }
        if (it >= maxit) { termcd = 4; break; }
        // forward-difference Jacobian: one batched gradient call for x + h_j e_j, j = 1..p
        std::vector<double> h(p);
        for (int j = 0; j < p; ++j) {
            h[j] = std::sqrt(2.220446049250313e-16) * std::max(std::fabs(x[j]), 1.0);
            for (int a = 0; a < p; ++a) B[(size_t)j * p + a] = x[a] + (a == j ? h[j] : 0.0);
        }
        MCL_TRY(F(B.data(), p, Fb.data()));
        for (int j = 0; j < p; ++j)
            for (int a = 0; a < p; ++a) J[(size_t)a * p + j] = (Fb[(size_t)j * p + a] - f[a]) / h[j];
        // solve J dx = -f (Gaussian elimination with partial pivoting; p <= MCLCAR_MAX_COV)
        std::vector<double> A(J);
        for (int a = 0; a < p; ++a) dx[a] = -f[a];
        bool singular = false;
        for (int c = 0; c < p && !singular; ++c) {
            int piv = c;
            for (int r = c + 1; r < p; ++r) if (std::fabs(A[(size_t)r * p + c]) > std::fabs(A[(size_t)piv * p + c])) piv = r;
            if (!(std::fabs(A[(size_t)piv * p + c]) > 0)) { singular = true; break; }
            if (piv != c) { for (int k = 0; k < p; ++k) std::swap(A[(size_t)c * p + k], A[(size_t)piv * p + k]); std::swap(dx[c], dx[piv]); }
            for (int r = c + 1; r < p; ++r) {
                const double m = A[(size_t)r * p + c] / A[(size_t)c * p + c];
                for (int k = c; k < p; ++k) A[(size_t)r * p + k] -= m * A[(size_t)c * p + k];
                dx[r] -= m * dx[c];
            }
        }
        if (singular) { termcd = 6; break; }   // nleqslv: Jacobian singular
        for (int c = p - 1; c >= 0; --c) {
            for (int k = c + 1; k < p; ++k) dx[c] -= A[(size_t)c * p + k] * dx[k];
            dx[c] /= A[(size_t)c * p + c];
        }
        const double f0 = sumsq(f);
        double t = 1.0;
        bool ok = false;
        for (int half = 0; half <= 4; ++half, t *= 0.5) {
            for (int a = 0; a < p; ++a) xn[a] = x[a] + t * dx[a];
            MCL_TRY(F(xn.data(), 1, fn.data()));
            if (sumsq(fn) < f0) { ok = true; break; }
        }
        ++it;
        if (!ok) { termcd = 3; break; }
        x = xn; f = fn;
    }
    for (int a = 0; a < p; ++a) beta_out[a] = x[a];
    if (fvec_out) for (int a = 0; a < p; ++a) fvec_out[a] = f[a];
    if (iters_out) *iters_out = it;
    if (termcd_out) *termcd_out = termcd;
    return MCLCAR_OK;
}

/* mcl.prep.glm(data, family, psi, ...) with postZ, R/mcl.glm.R:5-40: draws N latent fields
 * z | y at psi0 with the batched Metropolis-within-Gibbs sampler and attaches psi0. */
int mclcar_glm_prep(mclcar_ctx* ctx, int family, const double* psi0, int P, int64_t N, const mclcar_sampler_opts* opts,
                    mclcar_samples** out) {
    if (!ctx || !psi0 || !out) return MCLCAR_EINVAL;
    *out = nullptr;
    MCL_TRY(require_device(ctx));
    if (family != MCLCAR_FAMILY_BINOM && family != MCLCAR_FAMILY_POISSON) return fail(ctx, MCLCAR_EINVAL, "family must be binomial or Poisson");
    if (ctx->kind == GRAPH_NONE || !ctx->has_obs || ctx->p < 1) return fail(ctx, MCLCAR_ESTATE, "set graph, design (covX) and observations first");
    if (P != 2 + ctx->p) return fail(ctx, MCLCAR_EINVAL, "psi0 has length %d; expected %d", P, 2 + ctx->p);
    if (N < 1 || N >= ((int64_t)1 << 31)) return fail(ctx, MCLCAR_EINVAL, "bad sample count");
    MCL_TRY(need_eigs(ctx));
    const double rho = psi0[0], sigma = psi0[1];
    if (!(sigma > 0)) return fail(ctx, MCLCAR_EINVAL, "sigma must be positive");
    MCL_TRY(check_rho(ctx, rho));
    mclcar_sampler_opts o{};
    if (opts) o = *opts;
    o.keep = 1;
    if (o.dtype != MCLCAR_F64) o.dtype = MCLCAR_F32;
    int burn = o.burn, thin = o.thin;
    double omega = 1.0;  // Metropolis-within-Gibbs: no over-relaxation
    default_schedule(ctx, rho, &omega, &burn, &thin);
    // Proposals come from a Gaussian approximation of each site's full conditional (sampler_csr.cu, glm_laplace; every
    // third sweep from the prior conditional): most draws are accepted, and the data make the posterior field less
    // dependent than the prior one the Gibbs schedule is made for -- start from twice its burn-in and half its
    // thinning (measured on the config-3 map, scripts/exp_glm_proposal.py: effective sample size 0.94 N / 0.97 N at 6
    // sweeps per kept state for the binomial / Poisson model; proposals from the prior conditional alone,
    // MCLCAR_GLM_PROPOSAL=prior, are accepted 34 % / 43 % of the time and need 33 sweeps for 0.92 N).  Then CHECK: the
    // reference's v.lr and step rule take the draws as independent (R/mcl.bin.R:70-72), so the kept states must have an
    // effective sample size >= 0.85 N in both z'z and z'Wz; otherwise the run was a pilot (the reference's
    // mcl.prep.glm has one too, R/mcl.glm.R:17-26): its autocorrelation time gives the thinning that brings the
    // lag-one autocorrelation of kept states to 5 %, and the chains are run again.  A thinning given by the caller is
    // taken as is.
    const bool tune = o.thin <= 0;
    const char* pe = getenv("MCLCAR_GLM_PROPOSAL");
    const bool prior_proposal = pe && !strcmp(pe, "prior");
    if (o.burn <= 0) burn = prior_proposal ? std::max(4 * burn, 200) : std::max(2 * burn, 60);
    if (tune) thin = prior_proposal ? std::max(3 * thin, 10) : std::max((thin + 1) / 2 + 1, 4);
    const int n = ctx->n;
    const size_t es = o.dtype == MCLCAR_F32 ? 4 : 8;
    std::vector<double> xb(n, 0.0);
    for (int l = 0; l < ctx->p; ++l)
        for (int i = 0; i < n; ++i) xb[i] += ctx->X[(size_t)l * n + i] * psi0[2 + l];
    mclcar_samples* s = new (std::nothrow) mclcar_samples();
    if (!s) return fail(ctx, MCLCAR_ENOMEM, "out of host memory");
    s->ctx = ctx; s->dtype = o.dtype; s->sites = n; s->N = N; s->owned = true; s->layout = MCLCAR_ROW_IS_SAMPLE;
    const int64_t align = 16 / es;
    s->ld = (N + align - 1) / align * align;
    GmrfHost g;
    g.n = n; g.rowptr = &ctx->rowptr; g.colidx = &ctx->colidx; g.color_ptr = &ctx->color_ptr; g.order = &ctx->order;
    g.qdiag.assign(n, 1.0 / sigma);
    g.qoff.resize(ctx->colidx.size());
    for (size_t e = 0; e < g.qoff.size(); ++e) g.qoff[e] = -rho * (ctx->binary ? 1.0 : ctx->val[e]) / sigma;
    int st = pool_alloc(ctx, &s->d_M, (size_t)n * s->ld * es);
    const int64_t C = o.n_chains > 0 ? o.n_chains + (o.n_chains & 1) : default_csr_chains(ctx, n, N, burn, thin);
    const uint64_t draw0 = ctx->draw;
    int64_t sweeps_total = 0;
    // z'z and z'Wz of the kept states come out of the chain kernel itself (binary W, fp32 samples): at emission the state
    // and all its neighbours are in shared memory -- no second pass over the sample matrix (whose edge gathers miss L2 on
    // a map in arbitrary site order: 2.7 x the algorithmic traffic, profiles/r02_ncu_stats_site_major.md)
    if (st == MCLCAR_OK) st = samples_alloc_stats(ctx, s, 2);
    bool stats_done = false;
    for (int attempt = 0; attempt < 3 && st == MCLCAR_OK; ++attempt) {
        ctx->draw = draw0;               // one prep call = one draw index, whatever the number of pilot runs
        ctx->draw_salt = (uint64_t)attempt;
        st = run_csr_sampler(ctx, g, family == MCLCAR_FAMILY_BINOM ? FAM_BINOM : FAM_POIS, 1.0, nullptr, ctx->y.data(),
                             ctx->ntrial.empty() ? nullptr : ctx->ntrial.data(), xb.data(), N, C, burn, thin, s->d_M,
                             o.dtype == MCLCAR_F64, s->ld, &s->accept, &s->sweeps, getenv("MCLCAR_NO_CHAIN_STATS") ? nullptr : s->d_q, N,
                             &stats_done);
        sweeps_total += s->sweeps;
        s->stats_ready = stats_done;
        if (st != MCLCAR_OK || !tune || attempt == 2 || N < 4 * C) break;   // fewer than 4 states per chain: no autocorrelation to estimate
        // z'z, z'Wz of the kept states (they stay with the sample set) and their effective sample size
        if (!stats_done &&
            (st = launch_quadforms_site_major(ctx, s->d_M, s->dtype, N, s->ld, n, ctx->rowptr, ctx->colidx, ctx->val, 0, nullptr, s->d_q, N))) break;
        std::vector<double> q((size_t)2 * N);
        cudaError_t e = cudaMemcpyAsync(q.data(), s->d_q, q.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { st = fail(ctx, MCLCAR_ECUDA, "copy of the statistics failed: %s", cudaGetErrorString(e)); break; }
        double ess[2];
        mclcar_ess_from_stats(q.data(), N, 2, N, C, ess);
        const double worst = std::min(ess[0], ess[1]);
        if (worst >= 0.85 * (double)N) { s->stats_ready = true; break; }
        const double tau = (double)N / std::max(worst, 1.0);
        const double rho1 = std::min((tau - 1.0) / (tau + 1.0), 0.98);       // lag-one autocorrelation of the kept states
        const double per_sweep = std::pow(rho1, 1.0 / thin);
        int want = (int)std::ceil(std::log(0.05) / std::log(per_sweep));
        want = std::min(std::max(want, thin + thin / 4 + 1), 8 * thin);
        thin = want;
    }
    ctx->draw = draw0 + 1;
    ctx->draw_salt = 0;
    s->sweeps = sweeps_total;
    s->chains = C; s->burn = burn; s->thin = thin; s->omega = 1.0;
    if (st == MCLCAR_OK) st = mclcar_glm_attach(ctx, s, family, psi0, P, nullptr);
    if (st != MCLCAR_OK) {
        std::string keep_err = ctx->err;
        mclcar_samples_free(s);
        ctx->err = keep_err;
        return st;
    }
    *out = s;
    return MCLCAR_OK;
}

}  // extern "C"

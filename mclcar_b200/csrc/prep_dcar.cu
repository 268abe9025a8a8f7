// mcl.prep.dCAR (R/mcl.dCAR.R:2-8): draw N samples at the importance point psi0 and attach
// t10 = h(y; psi0), t20_i = h(Y_i; psi0).  The reference replicates CAR.simLM N times (one
// sparse Cholesky per sample); here N/C kept states of C parallel Gibbs chains.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <cstdlib>

#include "engine.hpp"

namespace mclcar {

// rho must keep I - rho W positive definite: (1/min(ew), 1/max(ew)), R/CAR.simLM.R:9-11
int check_rho(mclcar_ctx* ctx, double rho) {
    if (!ctx->eigs.empty() && ctx->ew_min < 0 && ctx->ew_max > 0) {
        const double lo = 1.0 / ctx->ew_min, hi = 1.0 / ctx->ew_max;
        if (rho < lo || rho > hi) return fail(ctx, MCLCAR_ERHO, "rho should be within %.10g %.10g", lo, hi);
        return MCLCAR_OK;
    }
    // no spectrum supplied: Gershgorin bound |rho| * max row sum < 1 is sufficient
    double rs = 0;
    for (int i = 0; i < ctx->n; ++i) {
        double a = 0;
        for (int e = ctx->rowptr[i]; e < ctx->rowptr[i + 1]; ++e) a += ctx->binary ? 1.0 : std::fabs(ctx->val[e]);
        rs = std::max(rs, a);
    }
    if (std::fabs(rho) * rs >= 1.0)
        return fail(ctx, MCLCAR_ERHO, "rho should be within %.10g %.10g (Gershgorin bound; supply the eigenvalues of W for the exact range)",
                    -1.0 / rs, 1.0 / rs);
    return MCLCAR_OK;
}

// Sweep schedule.  mu = |rho| * spectral radius of W is the Jacobi contraction factor.  Plain
// chromatic Gibbs (omega = 1) decorrelates at rate mu^2 per sweep; on a two-colourable graph
// (consistently ordered) the over-relaxed sampler with omega* = 2/(1 + sqrt(1 - mu^2)) does so
// at rate omega* - 1 (Young's SOR theory; Fox & Parker 2017 for samplers), with a linear
// polynomial factor at omega* itself (a Jordan block: k rate^k).  Slightly above omega* every
// eigenvalue of the sweep operator has modulus exactly omega - 1 and the factor is gone; what
// remains is the beat of the complex eigenvalue pair, measured on the torus at rho = 0.2 as a
// prefactor of <= 2.7 on the autocorrelation of the two slowest modes (constant field and
// checkerboard; scripts/acf_modes.py): lag 4 1.6 %, lag 5 0.3 % at omega* + 0.025 against 2.9 % and
// 0.8 % at omega*.  Default omega* + 0.025, prefactor 3 in the rule.
// thin: autocorrelation <= 1 %; burn: start bias <= 1e-6.
void default_schedule(const mclcar_ctx* ctx, double rho, double* omega, int* burn, int* thin) {
    double lam = 0;
    if (!ctx->eigs.empty()) lam = std::max(std::fabs(ctx->ew_min), std::fabs(ctx->ew_max));
    else
        for (int i = 0; i < ctx->n; ++i) {
            double a = 0;
            for (int e = ctx->rowptr[i]; e < ctx->rowptr[i + 1]; ++e) a += ctx->binary ? 1.0 : std::fabs(ctx->val[e]);
            lam = std::max(lam, a);
        }
    const double mu = std::min(std::max(std::fabs(rho) * lam, 1e-2), 0.9998);
    const double w_opt = 2.0 / (1.0 + std::sqrt(1.0 - mu * mu));
    if (!(*omega > 0)) *omega = ctx->ncolors == 2 ? w_opt + 0.025 : 1.0;
    if (*omega < 1.0) *omega = 1.0;
    if (*omega > 1.98) *omega = 1.98;
    const double w = *omega;
    double rate;
    bool poly = false;
    if (w >= w_opt - 1e-12) {
        rate = w - 1.0;
        poly = w > 1.0 && w < w_opt + 0.02;
    } else {
        const double t = 0.5 * (w * mu + std::sqrt(w * w * mu * mu - 4.0 * (w - 1.0)));
        rate = t * t;
    }
    rate = std::min(std::max(rate, 1e-3), 0.9996);
    auto sweeps_for = [&](double tol, int lo, int hi) {
        for (int k = lo; k < hi; ++k)
            if ((poly ? k : (w > 1.0 ? 3.0 : 1.0)) * std::pow(rate, k) <= tol) return k;
        return hi;
    };
    if (*thin <= 0) *thin = sweeps_for(0.01, 1, 4000);
    if (*burn <= 0) *burn = sweeps_for(1e-6, 4, 40000);
}

// Chains of the tiled torus sampler.  One sweep is one launch over all chains: its ramp-up (every
// CTA of the first wave waits for its tile) and tail cost about 9 us, so launches should carry
// >= 10^8 site updates (measured on B200: 33 chains of 10^6 sites reach 0.58 of the HBM roofline,
// 111+ chains 0.72).  More chains mean fewer kept states per chain, i.e. relatively more burn-in:
// keep burn-in below 1/8 of the kept sweeps, then pick the count in the upper fifth of the range
// with the smallest modelled time: launches (burn + ceil(N / C) thin) x (C + the launch overhead
// expressed in chains).
int64_t default_torus_chains(int n, int64_t N, int burn, int thin) {
    int64_t c_hi = std::max<int64_t>(1, ((int64_t)160 << 20) / n);
    c_hi = std::min<int64_t>(c_hi, std::max<int64_t>(1, N * thin / (8 * (int64_t)std::max(burn, 1))));
    c_hi = std::min<int64_t>(std::min<int64_t>(c_hi, N), 65535);
    const int64_t c_lo = std::max<int64_t>(1, c_hi - c_hi / 5);
    const double overhead = 5.4e6 / n;   // ~9 us of ramp-up and tail at ~6e11 site updates/s
    int64_t best = c_hi;
    double best_cost = 1e300;
    for (int64_t c = c_hi; c >= c_lo; --c) {
        const double cost = (double)(burn + (N + c - 1) / c * thin) * ((double)c + overhead);
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

// the tiled red-black kernels need an even x even torus; lattices whose chains (8 at a time) fit in
// shared memory go to the general chromatic kernel instead, which keeps them on chip for the whole run
bool use_torus_sampler(const mclcar_ctx* ctx) {
    return ctx->kind == GRAPH_TORUS && ctx->n1 % 2 == 0 && ctx->n2 % 2 == 0 && (size_t)ctx->n * 8 * sizeof(float) > 200 * 1024;
}

// Chains of the shared-memory chain sampler.  The upper end fills every SM's shared memory with chains; the schedule
// then decides how many of them pay: C chains cost C * burn sweeps before the first kept state and N * thin after, so
// the count is halved while the burn-in would still be more than half of the kept-state sweeps (never below two
// blocks of chains per SM).  Measured on one B200 (scripts/exp_chains.py): 10 x 10 HCAR map, 20 000 states, 18 944
// chains 0.43 ms -> 2 368 chains 0.24 ms; 40 x 40 binomial GLM, 10 000 states, 1 184 chains 16.8 ms -> 592 chains 15.5 ms.
int64_t default_csr_chains(const mclcar_ctx* ctx, int n, int64_t N, int burn, int thin) {
    const size_t cap = 200 * 1024;
    int cbmax = 0;
    for (int cb : {32, 16, 8, 4, 2})
        if ((size_t)(n + 1) * cb * sizeof(float) <= cap) { cbmax = cb; break; }
    if (!cbmax) return std::min<int64_t>(N + (N & 1), (int64_t)ctx->n_sm * 2 * 32);
    int per_sm = (int)std::min<size_t>(8, std::max<size_t>(1, cap / ((size_t)n * cbmax * sizeof(float))));
    int64_t c_auto = (int64_t)ctx->n_sm * per_sm * cbmax;
    if (burn > 0 && thin > 0 && !getenv("MCLCAR_CHAINS_FILL")) {
        const double c_pay = (double)N * thin / (2.0 * burn);
        while ((double)c_auto > c_pay && c_auto / 2 >= (int64_t)ctx->n_sm * 4) c_auto /= 2;
    }
    return std::max<int64_t>(2, std::min<int64_t>(c_auto, N + (N & 1)));
}

}  // namespace mclcar

using namespace mclcar;

extern "C" int mclcar_dcar_prep(mclcar_ctx* ctx, const double* psi0, int P, int64_t N, const mclcar_sampler_opts* opts,
                                mclcar_samples** out) {
    if (!ctx || !psi0 || !out) return MCLCAR_EINVAL;
    *out = nullptr;
    MCL_TRY(require_device(ctx));
    if (ctx->kind == GRAPH_NONE || !ctx->has_obs) return fail(ctx, MCLCAR_ESTATE, "set graph, design and observations first");
    if (P != 2 && P != 2 + ctx->p) return fail(ctx, MCLCAR_EINVAL, "psi0 has length %d; expected 2 or %d", P, 2 + ctx->p);
    if (N < 1 || N >= ((int64_t)1 << 31)) return fail(ctx, MCLCAR_EINVAL, "bad sample count %lld", (long long)N);
    const double rho = psi0[0], sigma = psi0[1];
    if (!(sigma > 0)) return fail(ctx, MCLCAR_EINVAL, "sigma must be positive");
    MCL_TRY(check_rho(ctx, rho));
    mclcar_sampler_opts o{};
    if (opts) o = *opts;
    else o.keep = 1;
    if (o.dtype != MCLCAR_F64) o.dtype = MCLCAR_F32;
    int burn = o.burn, thin = o.thin;
    double omega = o.omega;
    default_schedule(ctx, rho, &omega, &burn, &thin);
    const int n = ctx->n;
    const size_t es = o.dtype == MCLCAR_F32 ? 4 : 8;

    // mean of the sampled field, X beta0: one number when every design column is constant (known since the design
    // was set), else a vector -- formed and uploaded while the burn-in sweeps run, which do not need it
    const bool has_mu = P > 2;
    const bool mu_is_const = !has_mu || ctx->design_const;
    double mu0 = 0.0;
    if (has_mu)
        for (int l = 0; l < ctx->p; ++l) mu0 += ctx->X[(size_t)l * n] * psi0[2 + l];
    std::vector<double> mu;
    auto fill_mu = [&]() {
        mu.assign(n, 0.0);
        for (int l = 0; l < ctx->p; ++l)
            for (int i = 0; i < n; ++i) mu[i] += ctx->X[(size_t)l * n + i] * psi0[2 + l];
    };

    mclcar_samples* s = new (std::nothrow) mclcar_samples();
    if (!s) return fail(ctx, MCLCAR_ENOMEM, "out of host memory");
    s->ctx = ctx; s->dtype = o.dtype; s->sites = n; s->N = N; s->owned = true;
    int st = MCLCAR_OK;
    const bool use_torus = use_torus_sampler(ctx);
    double* d_mu = nullptr;
    if (use_torus) {
        // ---- spatially tiled red-black sampler, sample-major output
        int64_t C = o.n_chains > 0 ? o.n_chains : default_torus_chains(n, N, burn, thin);
        C = std::min<int64_t>(std::min<int64_t>(C, N), 65535);
        const int64_t E = (N + C - 1) / C;
        const int64_t align = 16 / es;
        const int64_t ld = (n + align - 1) / align * align;
        s->layout = MCLCAR_COL_IS_SAMPLE;
        s->ld = ld;
        TorusChains tc;
        void* staging = nullptr;
        begin_draw(ctx);   // a fresh Philox stream for every mcl.prep.dCAR call (R draws fresh rnorm() each time)
        do {
            // statistics-only sampling: the emitting sweep accumulates x'x, x'Wx and sum x of the samples it would
            // have written (no staging buffer, no statistics pass); otherwise samples go through a staging buffer of
            // one emission and the statistics kernel
            const bool fused_stats = o.dtype == MCLCAR_F32 && mu_is_const && ctx->design_const && torus_fused_tile_rows(ctx) > 0 &&
                                     !getenv("MCLCAR_NO_FUSED_STATS") && !(o.keep && getenv("MCLCAR_NO_KEPT_FUSED_STATS"));
            // kept samples under the same conditions: the emitting sweep writes the samples AND forms their statistics
            // from the registers it writes them from -- the sample matrix is not read back by a statistics pass
            if (o.keep || !fused_stats)
                if ((st = pool_alloc(ctx, o.keep ? &s->d_M : &staging, (size_t)(o.keep ? N : C) * ld * es))) break;
            if ((st = torus_chains_alloc(ctx, C, tc))) break;
            if ((!o.keep || fused_stats) && (st = samples_alloc_stats(ctx, s, 2 + 2 * ctx->p))) break;
            const bool fused = (ctx->n2 / 2) % 4 == 0;   // emission fused into the last black half sweep
            uint32_t sweep = 0;
            // sweeps that emit nothing go two to a launch where the lattice allows (the tile stays in shared memory)
            auto plain_sweeps = [&](int count) {
                int rc = MCLCAR_OK;
                while (count > 0 && rc == MCLCAR_OK) {
                    int made = 1;
                    rc = torus_sweep(ctx, tc, rho, sigma, omega, sweep, nullptr, count, &made);
                    sweep += (uint32_t)made;
                    count -= made;
                }
                return rc;
            };
            st = plain_sweeps(burn);
            if (st == MCLCAR_OK && !mu_is_const) {
                if ((st = pool_alloc(ctx, (void**)&d_mu, (size_t)n * sizeof(double)))) break;
                fill_mu();
                // fp32 samples: the emitting kernels add the mean as an fp32 pair (hi, lo) stored in the 8-byte slot
                if (o.dtype == MCLCAR_F32)
                    for (int i = 0; i < n; ++i) {
                        const float hi = (float)mu[i], lo = (float)(mu[i] - (double)hi);
                        float pair[2] = {hi, lo};
                        std::memcpy(&mu[i], pair, sizeof(double));
                    }
                cudaError_t e = cudaMemcpyAsync(d_mu, mu.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // mu is a local
                if (e != cudaSuccess) { st = fail(ctx, MCLCAR_ECUDA, "upload of X beta0 failed: %s", cudaGetErrorString(e)); break; }
            }
            for (int64_t e_i = 0; e_i < E && st == MCLCAR_OK; ++e_i) {
                const int64_t nch = std::min<int64_t>(C, N - e_i * C);
                TorusEmit em;
                em.out = o.keep ? s->d_M : staging; em.dtype = o.dtype; em.ld = ld; em.slot0 = o.keep ? e_i * C : 0; em.nch = nch;
                em.d_mu = mu_is_const ? nullptr : d_mu;
                em.mu_const = mu_is_const ? mu0 : 0.0;
                if (fused_stats) {
                    if (!o.keep) em.out = nullptr;
                    em.slot0 = e_i * C; em.stats_q = s->d_q; em.stats_ldq = N;
                }
                if (fused) {
                    st = plain_sweeps(thin - 1);
                    if (st == MCLCAR_OK) st = torus_sweep(ctx, tc, rho, sigma, omega, sweep++, &em);
                } else {
                    st = plain_sweeps(thin);
                }
                if (st) break;
                if (fused_stats) continue;
                if (!fused) st = torus_emit(ctx, tc, em.d_mu, em.mu_const, em.out, o.dtype, ld, em.slot0, nch);
                if (st == MCLCAR_OK && !o.keep) {
                    bool handled = false;
                    st = launch_stats_torus_fast(ctx, staging, o.dtype, nch, ld, s->d_q + e_i * C, N, &handled);
                    if (st == MCLCAR_OK && !handled) st = fail(ctx, MCLCAR_ELIMIT, "statistics-only sampling needs n2 %% 4 == 0 and rows <= 24 KB");
                }
            }
            s->sweeps = sweep;
            s->chains = C;
            // every launch of the run is queued: the host forms the algebra of a freshly set design / observations
            // (W X, X'X, the statistics of y) now, while the device sweeps
            if (st == MCLCAR_OK) st = ensure_algebra(ctx);
            if (st == MCLCAR_OK && (!o.keep || fused_stats)) {
                if ((st = samples_update_qbar(ctx, s))) break;
                s->stats_ready = true;
            }
        } while (0);
        torus_chains_free(ctx, tc);
        pool_free(ctx, staging);
    } else {
        // ---- generic chromatic sampler with shared-memory-resident chains, site-major output
        s->layout = MCLCAR_ROW_IS_SAMPLE;
        const int64_t align = 16 / es;
        const int64_t ld = (N + align - 1) / align * align;
        s->ld = ld;
        GmrfHost g;
        g.n = n; g.rowptr = &ctx->rowptr; g.colidx = &ctx->colidx; g.color_ptr = &ctx->color_ptr; g.order = &ctx->order;
        g.qdiag.assign(n, 1.0 / sigma);
        g.qoff.resize(ctx->colidx.size());
        for (size_t e = 0; e < g.qoff.size(); ++e) g.qoff[e] = -rho * (ctx->binary ? 1.0 : ctx->val[e]) / sigma;
        st = pool_alloc(ctx, &s->d_M, (size_t)n * ld * es);
        const int64_t C = o.n_chains > 0 ? o.n_chains + (o.n_chains & 1) : default_csr_chains(ctx, n, N, burn, thin);
        if (has_mu) fill_mu();
        if (st == MCLCAR_OK)
            st = run_csr_sampler(ctx, g, FAM_GAUSS, omega, mu.empty() ? nullptr : mu.data(), nullptr, nullptr, nullptr, N, C, burn, thin,
                                 s->d_M, o.dtype == MCLCAR_F64, ld, &s->accept, &s->sweeps);
        s->chains = C;
    }
    pool_free(ctx, d_mu);
    s->burn = burn; s->thin = thin; s->omega = omega;
    if (st == MCLCAR_OK) st = mclcar_dcar_attach(ctx, s, psi0, P, nullptr);
    if (st == MCLCAR_OK && !o.keep && s->d_M) {  // statistics only on the generic path: compute, then drop the matrix
        st = dcar_ensure_stats(ctx, s);
        pool_free(ctx, s->d_M);
        s->d_M = nullptr;
    }
    if (st != MCLCAR_OK) {
        std::string keep_err = ctx->err;
        mclcar_samples_free(s);
        ctx->err = keep_err;
        return st;
    }
    *out = s;
    return MCLCAR_OK;
}

/* What mclcar_dcar_prep would run for (rho, N, opts) on this context's graph, without running it:
 * sampler kind (1 = tiled red-black torus kernels, 2 = general chromatic kernel with the chains in
 * shared memory, 3 = same with the chains in global memory), parallel chains, burn-in sweeps, sweeps
 * per kept state, over-relaxation factor.  Pure host logic: also answers on a host-only context. */
extern "C" int mclcar_dcar_plan(mclcar_ctx* ctx, double rho, int64_t N, const mclcar_sampler_opts* opts, int* kind,
                                int64_t* chains, int* burn_out, int* thin_out, double* omega_out) {
    if (!ctx) return MCLCAR_EINVAL;
    if (ctx->kind == GRAPH_NONE) return fail(ctx, MCLCAR_ESTATE, "set the graph first");
    if (N < 1 || N >= ((int64_t)1 << 31)) return fail(ctx, MCLCAR_EINVAL, "bad sample count %lld", (long long)N);
    MCL_TRY(check_rho(ctx, rho));
    mclcar_sampler_opts o{};
    if (opts) o = *opts;
    int burn = o.burn, thin = o.thin;
    double omega = o.omega;
    default_schedule(ctx, rho, &omega, &burn, &thin);
    const int n = ctx->n;
    const bool use_torus = use_torus_sampler(ctx);
    int64_t C;
    int k;
    if (use_torus) {
        C = o.n_chains > 0 ? o.n_chains : default_torus_chains(n, N, burn, thin);
        C = std::min<int64_t>(std::min<int64_t>(C, N), 65535);
        k = 1;
    } else {
        C = o.n_chains > 0 ? o.n_chains + (o.n_chains & 1) : default_csr_chains(ctx, n, N, burn, thin);
        k = (size_t)(n + 1) * 2 * sizeof(float) <= 200 * 1024 ? 2 : 3;
    }
    if (kind) *kind = k;
    if (chains) *chains = C;
    if (burn_out) *burn_out = burn;
    if (thin_out) *thin_out = thin;
    if (omega_out) *omega_out = omega;
    return MCLCAR_OK;
}

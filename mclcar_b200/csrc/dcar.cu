// Direct (linear-model) CAR: host logic around the statistics (K2) and reduction (K3)
// kernels.  Mirrors R/mcl.dCAR.R of the reference function by function.
//
// With q = (S1, S2, a[p], b[p]) = (Y'Y, Y'WY, X'Y, X'WY) and G0 = X'X, G1 = X'WX:
//   r'r  = S1 - 2 beta'a + beta'G0 beta,     r'Wr = S2 - 2 beta'b + beta'G1 beta,
//   X'r  = a - G0 beta,                       X'Wr = b - G1 beta,
//   h(Y; psi) = -(r'r - rho r'Wr)/(2 sigma)                          (logunnorm, :168-185)
//   D.log.unorm = ( r'Wr/(2 sigma), (r'r - rho r'Wr)/(2 sigma^2), (X'r - rho X'Wr)/sigma )  (:188-210)
// all affine in q for a fixed candidate psi -- see SURVEY.md section 8.
#include <cmath>
#include <cstring>

#include "engine.hpp"

namespace mclcar {

namespace {

// affine maps of one candidate: h = c0 + c.q ; features f_a = a0[a] + A[a].q ; X'Wr_l = xw0[l] + q[2+p+l]
struct DcarAffine {
    int p = 0, d = 0, P = 0;
    double c0 = 0;
    std::vector<double> c;       // d
    std::vector<double> a0, A;   // P, P x d
    std::vector<double> xw0;     // p: -(G1 beta)_l
    double rho = 0, sigma = 1;
};

int build_affine(mclcar_ctx* ctx, const double* pars, int P, DcarAffine& m) {
    MCL_TRY(ensure_algebra(ctx));
    const int p = ctx->p, d = 2 + 2 * p;
    if (P != 2 && P != 2 + p) return fail(ctx, MCLCAR_EINVAL, "parameter vector of length %d; expected 2 or %d", P, 2 + p);
    const int pb = P - 2;  // regression coefficients actually used
    const double rho = pars[0], sigma = pars[1];
    if (!(sigma > 0) || !std::isfinite(rho)) return fail(ctx, MCLCAR_EINVAL, "need sigma > 0 and finite rho (got %g, %g)", sigma, rho);
    m.p = p; m.d = d; m.P = P; m.rho = rho; m.sigma = sigma;
    const double* beta = pars + 2;
    std::vector<double> G0b(p, 0.0), G1b(p, 0.0);
    double bG0b = 0, bG1b = 0;
    for (int l = 0; l < pb; ++l) {
        for (int k = 0; k < pb; ++k) {
            G0b[l] += ctx->G0[(size_t)l * p + k] * beta[k];
            G1b[l] += ctx->G1[(size_t)l * p + k] * beta[k];
        }
        bG0b += beta[l] * G0b[l];
        bG1b += beta[l] * G1b[l];
    }
    m.c.assign(d, 0.0);
    m.c0 = -(bG0b - rho * bG1b) / (2 * sigma);
    m.c[0] = -1.0 / (2 * sigma);
    m.c[1] = rho / (2 * sigma);
    for (int l = 0; l < pb; ++l) {
        m.c[2 + l] = beta[l] / sigma;
        m.c[2 + p + l] = -rho * beta[l] / sigma;
    }
    m.a0.assign(P, 0.0);
    m.A.assign((size_t)P * d, 0.0);
    // d/d rho = r'Wr / (2 sigma)
    m.a0[0] = bG1b / (2 * sigma);
    m.A[1] = 1.0 / (2 * sigma);
    for (int l = 0; l < pb; ++l) m.A[2 + p + l] = -beta[l] / sigma;
    // d/d sigma = (r'r - rho r'Wr)/(2 sigma^2) = -h / sigma
    m.a0[1] = -m.c0 / sigma;
    for (int j = 0; j < d; ++j) m.A[(size_t)1 * d + j] = -m.c[j] / sigma;
    // d/d beta_l = (X'r - rho X'Wr)_l / sigma
    for (int l = 0; l < pb; ++l) {
        m.a0[2 + l] = (-G0b[l] + rho * G1b[l]) / sigma;
        m.A[(size_t)(2 + l) * d + 2 + l] = 1.0 / sigma;
        m.A[(size_t)(2 + l) * d + 2 + p + l] = -rho / sigma;
    }
    m.xw0.assign(p, 0.0);
    for (int l = 0; l < pb; ++l) m.xw0[l] = -G1b[l];
    return MCLCAR_OK;
}

double affine_h(const DcarAffine& m, const double* q) {
    double h = m.c0;
    for (int j = 0; j < m.d; ++j) h += m.c[j] * q[j];
    return h;
}
void affine_g(const DcarAffine& m, const double* q, double* g) {
    for (int a = 0; a < m.P; ++a) {
        double f = m.a0[a];
        for (int j = 0; j < m.d; ++j) f += m.A[(size_t)a * m.d + j] * q[j];
        g[a] = f;
    }
}

// second derivatives of h from (g, X'Wr): R/mcl.dCAR.R:331-367
void d2_from(const mclcar_ctx* ctx, const DcarAffine& m, const double* g, const double* xwr, double* D2) {
    const int P = m.P, p = m.p, pb = P - 2;
    const double sigma = m.sigma, rho = m.rho;
    for (int t = 0; t < P * P; ++t) D2[t] = 0.0;
    D2[1 * P + 1] = -2.0 * g[1] / sigma;
    D2[0 * P + 1] = D2[1 * P + 0] = -g[0] / sigma;
    for (int l = 0; l < pb; ++l) {
        D2[0 * P + 2 + l] = D2[(2 + l) * P + 0] = -xwr[l] / sigma;
        D2[1 * P + 2 + l] = D2[(2 + l) * P + 1] = -g[2 + l] / sigma;
        for (int k = 0; k < pb; ++k)
            D2[(2 + l) * P + 2 + k] = -(ctx->G0[(size_t)l * p + k] - rho * ctx->G1[(size_t)l * p + k]) / sigma;
    }
}

int check_ready(mclcar_ctx* ctx, mclcar_samples* s, bool need_attach) {
    if (!ctx || !s) return MCLCAR_EINVAL;
    if (s->ctx != ctx) return fail(ctx, MCLCAR_EINVAL, "sample set belongs to another context");
    if (ctx->kind == GRAPH_NONE || !ctx->has_obs) return fail(ctx, MCLCAR_ESTATE, "set graph, design and observations first");
    if (s->sites != ctx->n) return fail(ctx, MCLCAR_EINVAL, "sample matrix has %lld sites, graph has %d", (long long)s->sites, ctx->n);
    if (need_attach && !s->attached) return fail(ctx, MCLCAR_ESTATE, "call mclcar_dcar_attach (or _prep) first");
    return MCLCAR_OK;
}

}  // namespace

int dcar_ensure_stats(mclcar_ctx* ctx, mclcar_samples* s) {
    const int d = 2 + 2 * ctx->p;
    if (s->stats_ready && s->d == d) return MCLCAR_OK;
    if (!s->d_M) return fail(ctx, MCLCAR_ESTATE, "sample set has neither a matrix nor statistics");
    MCL_TRY(samples_alloc_stats(ctx, s, d));
    bool handled = false;
    if (s->layout == MCLCAR_COL_IS_SAMPLE) {
        if (ctx->kind == GRAPH_TORUS)
            MCL_TRY(launch_stats_torus_fast(ctx, s->d_M, s->dtype, s->N, s->ld, s->d_q, s->N, &handled));
        if (!handled) {
            // general graph, sample-major: transpose to site-major on the device, then the lane-per-sample kernel
            const size_t es = s->dtype == MCLCAR_F32 ? 4 : 8;
            void* t = nullptr;
            MCL_TRY(pool_alloc(ctx, &t, (size_t)s->sites * s->N * es));
            int st = launch_transpose(ctx, s->d_M, s->dtype, s->N, s->sites, s->ld, t, s->N);
            if (st == MCLCAR_OK) st = launch_stats_site_major(ctx, t, s->dtype, s->N, s->N, s->d_q, s->N);
            pool_free(ctx, t);
            MCL_TRY(st);
        }
    } else {
        MCL_TRY(launch_stats_site_major(ctx, s->d_M, s->dtype, s->N, s->ld, s->d_q, s->N));
    }
    MCL_TRY(samples_update_qbar(ctx, s));
    s->stats_ready = true;
    return MCLCAR_OK;
}

// coefficient blocks for the K candidates of a batch
static int build_plan(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int what, int shift_mode,
                      const int64_t* subset_idx, const int64_t* subset_ptr, ReducePlan& plan) {
    const int d = s->d;
    const int F = what == WHAT_LR ? 0 : P;
    plan.d_q = s->d_q; plan.ldq = s->N; plan.d = d; plan.N = s->N; plan.K = K; plan.F = F; plan.what = what;
    plan.shift_mode = shift_mode;
    plan.d_base = s->d_base;
    plan.subset_idx = subset_idx; plan.subset_ptr = subset_ptr;
    const int cs = coef_stride_of(d, F);
    plan.coef_stride = cs;
    plan.coef.assign((size_t)K * cs, 0.0);
    DcarAffine m0, m;
    MCL_TRY(build_affine(ctx, s->psi0.data(), (int)s->psi0.size(), m0));
    for (int k = 0; k < K; ++k) {
        MCL_TRY(build_affine(ctx, psi + (size_t)k * P, P, m));
        double* cf = plan.coef.data() + (size_t)k * cs;
        if (s->d_base) {  // the reference's own t20 was supplied: l = h(q; psi_k) - t20
            cf[0] = m.c0;
            for (int j = 0; j < d; ++j) cf[1 + j] = m.c[j];
        } else {          // l = h(q; psi_k) - h(q; psi0), coefficients differenced first
            cf[0] = m.c0 - m0.c0;
            for (int j = 0; j < d; ++j) cf[1 + j] = m.c[j] - m0.c[j];
        }
        double shift = cf[0];
        for (int j = 0; j < d; ++j) shift += cf[1 + j] * s->qbar[j];
        if (s->d_base) shift -= s->base_mean;
        cf[coef_off_shift(d)] = shift;
        int64_t ns = s->N;
        if (subset_ptr && subset_ptr[k + 1] > subset_ptr[k]) ns = subset_ptr[k + 1] - subset_ptr[k];
        cf[coef_off_n(d)] = (double)ns;
        cf[coef_off_centre(d)] = 1.0;
        for (int a = 0; a < F; ++a) {
            double* fa = cf + coef_off_feat(d, F) + (size_t)a * (1 + d);
            fa[0] = m.a0[a];
            double fbar = m.a0[a];
            for (int j = 0; j < d; ++j) {
                fa[1 + j] = m.A[(size_t)a * d + j];
                fbar += fa[1 + j] * s->qbar[j];
            }
            cf[coef_off_centre(d) + 1 + a] = fbar;
        }
    }
    return MCLCAR_OK;
}

int dcar_partial(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int what,
                 const int64_t* subset_idx, const int64_t* subset_ptr, int shift_mode, double* tuples, bool merge_ranks) {
    MCL_TRY(check_ready(ctx, s, true));
    MCL_TRY(require_device(ctx));   // binds this host thread to the context's device
    if (K < 1 || !psi || !tuples) return fail(ctx, MCLCAR_EINVAL, "empty candidate batch");
    if (P > MCLCAR_MAX_P) return fail(ctx, MCLCAR_ELIMIT, "P = %d exceeds MCLCAR_MAX_P", P);
    if (what < WHAT_LR || what > WHAT_HESS) return fail(ctx, MCLCAR_EINVAL, "bad 'what' %d", what);
    MCL_TRY(dcar_ensure_stats(ctx, s));
    ReducePlan plan;
    MCL_TRY(build_plan(ctx, s, psi, K, P, what, shift_mode, subset_idx, subset_ptr, plan));
    return run_reduce(ctx, plan, tuples, merge_ranks);
}

// tuples -> reference outputs (host only)
int dcar_finish(mclcar_ctx* ctx, const double* psi0, int P0, const double* psi, int K, int P, int what, int evar,
                const double* rho_cons, const double* tuples, double* out0, double* out1) {
    if (!ctx || !psi0 || !psi || !tuples || !out0) return MCLCAR_EINVAL;
    if (!ctx->has_obs) return fail(ctx, MCLCAR_ESTATE, "set graph, design and observations first");
    MCL_TRY(ensure_algebra(ctx));
    const int d = 2 + 2 * ctx->p;
    const int F = what == WHAT_LR ? 0 : P;
    const int64_t T = tuple_len(what, F, d);
    DcarAffine m0, m;
    MCL_TRY(build_affine(ctx, psi0, P0, m0));
    const double t10 = affine_h(m0, ctx->qobs.data());
    std::vector<double> vbar(1 + F), C((size_t)(1 + F) * (1 + F)), g10(P), gy(P), D2y((size_t)P * P),
        gw(P), A((size_t)P * P), xwr(ctx->p), ga(P);
    for (int k = 0; k < K; ++k) {
        const double* pars = psi + (size_t)k * P;
        const double* tup = tuples + (size_t)k * T;
        MCL_TRY(build_affine(ctx, pars, P, m));
        const double n = tup[0], shift = tup[1];
        if (what == WHAT_LR) {
            const bool inside = !rho_cons || (pars[0] > rho_cons[0] && pars[0] < rho_cons[1]);
            if (!inside) {  // R/mcl.dCAR.R:39-42
                out0[k] = -1e8;
                if (out1) out1[k] = 1e8;
                continue;
            }
            centred_moments(tup, 1, vbar.data(), C.data());
            const double s1 = affine_h(m, ctx->qobs.data()) - t10;
            out0[k] = s1 - (shift + std::log(vbar[0]));
            if (out1) out1[k] = C[0] / (n - 1.0) / (vbar[0] * vbar[0]) / n;
        } else if (what == WHAT_GRAD) {
            const int V = 1 + P;
            centred_moments(tup, V, vbar.data(), C.data());
            affine_g(m, ctx->qobs.data(), g10.data());
            const double eb = vbar[0];
            for (int a = 0; a < P; ++a) out0[(size_t)k * P + a] = g10[a] - vbar[1 + a] / eb;
            if (out1 && evar != 0) {
                double* o = out1 + (size_t)k * P * P;
                for (int a = 0; a < P; ++a)
                    for (int b = 0; b < P; ++b) {
                        double v = C[(size_t)(1 + a) * V + 1 + b];
                        if (evar >= 2)  // weights times (g - g10): R/mcl.dCAR.R:277-279 and MCMLE.var :163
                            v += -g10[a] * C[(size_t)0 * V + 1 + b] - C[(size_t)(1 + a) * V + 0] * g10[b] + g10[a] * g10[b] * C[0];
                        if (evar == 3) o[a * P + b] = std::exp(2.0 * shift) * v / (n - 1.0);
                        else o[a * P + b] = v / (n - 1.0) / (eb * eb) / n;
                    }
            }
        } else {  // Hessian, R/mcl.dCAR.R:369-393
            const int p = ctx->p, pb = P - 2;
            const double Se = tup[2];
            const double* Sf = tup + 3;
            const double* Sff = tup + 3 + P;
            const double* R1 = tup + 3 + P + P * (P + 1) / 2;
            for (int a = 0; a < P; ++a) gw[a] = Sf[a] / Se;
            for (int l = 0; l < pb; ++l) xwr[l] = R1[2 + p + l] / Se + m.xw0[l];
            d2_from(ctx, m, gw.data(), xwr.data(), A.data());  // mean(w D2h(Y_i))
            affine_g(m, ctx->qobs.data(), gy.data());
            for (int l = 0; l < pb; ++l) xwr[l] = ctx->qobs[2 + p + l] + m.xw0[l];
            d2_from(ctx, m, gy.data(), xwr.data(), D2y.data());
            double* o = out0 + (size_t)k * P * P;
            int sidx = 0;
            std::vector<double> B((size_t)P * P);
            for (int a = 0; a < P; ++a)
                for (int b = a; b < P; ++b, ++sidx) B[a * P + b] = B[b * P + a] = Sff[sidx] / Se;
            for (int a = 0; a < P; ++a)
                for (int b = 0; b < P; ++b) o[a * P + b] = D2y[a * P + b] - (A[a * P + b] + B[a * P + b] - gw[a] * gw[b]);
        }
    }
    return MCLCAR_OK;
}

static int dcar_run(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int what, int evar,
                    const double* rho_cons, const int64_t* subset_idx, const int64_t* subset_ptr, int shift_mode,
                    double* out0, double* out1) {
    MCL_TRY(check_ready(ctx, s, true));
    const int d = 2 + 2 * ctx->p;
    const int F = what == WHAT_LR ? 0 : P;
    const int64_t T = tuple_len(what, F, d);
    if (T < 0) return fail(ctx, MCLCAR_EINVAL, "bad 'what'");
    std::vector<double> tuples((size_t)K * T);
    MCL_TRY(dcar_partial(ctx, s, psi, K, P, what, subset_idx, subset_ptr, shift_mode, tuples.data(), true));
    MCL_TRY(gather_and_merge(ctx, what, F, d, K, tuples.data()));
    return dcar_finish(ctx, s->psi0.data(), (int)s->psi0.size(), psi, K, P, what, evar, rho_cons, tuples.data(), out0, out1);
}

}  // namespace mclcar

using namespace mclcar;

extern "C" {

int mclcar_dcar_stats(mclcar_ctx* ctx, mclcar_samples* s, double* q_out) {
    MCL_TRY(check_ready(ctx, s, false));
    MCL_TRY(require_device(ctx));
    MCL_TRY(dcar_ensure_stats(ctx, s));
    if (q_out) {
        MCL_CUDA(ctx, cudaMemcpyAsync(q_out, s->d_q, (size_t)s->d * s->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MCLCAR_OK;
}

int mclcar_dcar_attach(mclcar_ctx* ctx, mclcar_samples* s, const double* psi0, int P, const double* t20) {
    MCL_TRY(check_ready(ctx, s, false));
    MCL_TRY(require_device(ctx));
    if (!psi0) return MCLCAR_EINVAL;
    MCL_TRY(ensure_algebra(ctx));
    DcarAffine m0;
    MCL_TRY(build_affine(ctx, psi0, P, m0));
    s->psi0.assign(psi0, psi0 + P);
    s->t10 = affine_h(m0, ctx->qobs.data());
    if (s->d_base) {
        pool_free(ctx, s->d_base);
        s->d_base = nullptr;
    }
    s->base_mean = 0.0;
    if (t20) {
        MCL_TRY(pool_alloc(ctx, (void**)&s->d_base, (size_t)s->N * sizeof(double)));
        MCL_CUDA(ctx, cudaMemcpyAsync(s->d_base, t20, (size_t)s->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        long double a = 0;
        for (int64_t i = 0; i < s->N; ++i) a += t20[i];
        s->base_mean = (double)(a / s->N);
    }
    s->attached = true;
    return MCLCAR_OK;
}

int mclcar_dcar_t10_t20(mclcar_ctx* ctx, mclcar_samples* s, double* t10, double* t20) {
    MCL_TRY(check_ready(ctx, s, true));
    MCL_TRY(require_device(ctx));
    if (t10) *t10 = s->t10;
    if (t20) {
        if (s->d_base) {
            MCL_CUDA(ctx, cudaMemcpyAsync(t20, s->d_base, (size_t)s->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        } else {
            MCL_TRY(dcar_ensure_stats(ctx, s));
            std::vector<double> q((size_t)s->d * s->N);
            MCL_CUDA(ctx, cudaMemcpyAsync(q.data(), s->d_q, q.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            DcarAffine m0;
            MCL_TRY(build_affine(ctx, s->psi0.data(), (int)s->psi0.size(), m0));
            std::vector<double> qi(s->d);
            for (int64_t i = 0; i < s->N; ++i) {
                for (int j = 0; j < s->d; ++j) qi[j] = q[(size_t)j * s->N + i];
                t20[i] = affine_h(m0, qi.data());
            }
        }
    }
    return MCLCAR_OK;
}

int mclcar_dcar_eval(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, const double* rho_cons,
                     const int64_t* subset_idx, const int64_t* subset_ptr, int shift_mode, double* lr, double* var) {
    if (!lr) return MCLCAR_EINVAL;
    return dcar_run(ctx, s, psi, K, P, WHAT_LR, 0, rho_cons, subset_idx, subset_ptr, shift_mode, lr, var);
}

int mclcar_dcar_grad(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int evar, int shift_mode,
                     double* grad, double* gradvar) {
    if (!grad || evar < 0 || evar > 2) return MCLCAR_EINVAL;
    return dcar_run(ctx, s, psi, K, P, WHAT_GRAD, evar, nullptr, nullptr, nullptr, shift_mode, grad, gradvar);
}

int mclcar_dcar_hess(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int shift_mode, double* hess) {
    if (!hess) return MCLCAR_EINVAL;
    return dcar_run(ctx, s, psi, K, P, WHAT_HESS, 0, nullptr, nullptr, nullptr, shift_mode, hess, nullptr);
}

int mclcar_dcar_mcmle_var(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, double* out) {
    if (!out || K < 1) return MCLCAR_EINVAL;
    std::vector<double> grad((size_t)K * P);
    return dcar_run(ctx, s, psi, K, P, WHAT_GRAD, 3, nullptr, nullptr, nullptr, MCLCAR_SHIFT_MEAN, grad.data(), out);
}

int mclcar_dcar_partial(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int what,
                        const int64_t* subset_idx, const int64_t* subset_ptr, int shift_mode, double* tuples) {
    return dcar_partial(ctx, s, psi, K, P, what, subset_idx, subset_ptr, shift_mode, tuples, false);
}

int mclcar_dcar_finish(mclcar_ctx* ctx, const double* psi0, int P0, const double* psi, int K, int P, int what, int evar,
                       const double* rho_cons, const double* tuples, double* out0, double* out1) {
    return dcar_finish(ctx, psi0, P0, psi, K, P, what, evar, rho_cons, tuples, out0, out1);
}

}  // extern "C"

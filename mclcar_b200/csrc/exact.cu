// "Next" rows of the scope table: the profile / GLS step and the
// exact direct-CAR likelihood, evaluated for whole batches from the p x p design constants and
// the spectrum of W instead of the reference's dense n x n products.
//   sigmabeta(rho)      R/loglik.dCAR.R:72-86   beta = (X'Q1 X)^-1 X'Q1 y, sigma = r'Q1 r / n, Q1 = I - rho W
//   mcl.profile.dCAR    R/mcl.dCAR.R:53-96      mcl.dCAR at (rho, sigmabeta(rho))
//   loglik.dCAR         R/loglik.dCAR.R:7-42    -n/2 (log 2pi + log sigma) + 1/2 sum log(1 - rho lam) - r'Q1 r/(2 sigma)
//   ploglik.dCAR        R/loglik.dCAR.R:45-69   -n/2 log sigma_hat(rho) + 1/2 sum log(1 - rho lam) - n/2
// The log-determinant sum over the n eigenvalues for K values of rho runs on the device
// (logdet_kernel: one CTA per rho, fp64 log1p, fixed-order tree).
#include <cmath>

#include "common.cuh"
#include "engine.hpp"

namespace mclcar {
namespace {

// sum_i w_i log(1 - rho lam_i) over the spectral nodes (w null: every node counts once)
__global__ void __launch_bounds__(256) logdet_kernel(const double* eig, const double* w, int n, const double* rho, double* out) {
    __shared__ double scratch[8];
    const double r = rho[blockIdx.x];
    double a[1] = {0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) a[0] += (w ? w[i] : 1.0) * log1p(-r * eig[i]);
    block_sum<1>(a, scratch);
    if (threadIdx.x == 0) out[blockIdx.x] = a[0];
}

// solve the p x p system A x = b (A symmetric positive definite here) by Gaussian elimination with pivoting
int solve_small(std::vector<double> A, std::vector<double> b, int p, double* x) {
    for (int c = 0; c < p; ++c) {
        int piv = c;
        for (int r = c + 1; r < p; ++r)
            if (std::fabs(A[r * p + c]) > std::fabs(A[piv * p + c])) piv = r;
        if (A[piv * p + c] == 0.0) return 1;
        if (piv != c) {
            for (int k = 0; k < p; ++k) std::swap(A[c * p + k], A[piv * p + k]);
            std::swap(b[c], b[piv]);
        }
        for (int r = c + 1; r < p; ++r) {
            const double f = A[r * p + c] / A[c * p + c];
            for (int k = c; k < p; ++k) A[r * p + k] -= f * A[c * p + k];
            b[r] -= f * b[c];
        }
    }
    for (int r = p - 1; r >= 0; --r) {
        double s = b[r];
        for (int k = r + 1; k < p; ++k) s -= A[r * p + k] * x[k];
        x[r] = s / A[r * p + r];
    }
    return 0;
}

// (sigma, beta) at one rho from G0, G1 and the observed-data statistics
int sigmabeta_one(mclcar_ctx* ctx, double rho, double* out) {
    MCL_TRY(ensure_algebra(ctx));
    const int p = ctx->p, n = ctx->n;
    const double* q = ctx->qobs.data();
    std::vector<double> A((size_t)p * p), b(p), beta(p);
    for (int a = 0; a < p; ++a) {
        for (int c = 0; c < p; ++c) A[a * p + c] = ctx->G0[(size_t)a * p + c] - rho * ctx->G1[(size_t)a * p + c];
        b[a] = q[2 + a] - rho * q[2 + p + a];
    }
    if (p > 0 && solve_small(A, b, p, beta.data())) return fail(ctx, MCLCAR_EINVAL, "X'(I - rho W)X is singular at rho = %g", rho);
    // r'Q1 r = y'Q1 y - 2 beta'X'Q1 y + beta'X'Q1 X beta
    double rq = q[0] - rho * q[1];
    for (int a = 0; a < p; ++a) {
        rq -= 2.0 * beta[a] * b[a];
        for (int c = 0; c < p; ++c) rq += beta[a] * A[a * p + c] * beta[c];
    }
    out[0] = rq / n;
    for (int a = 0; a < p; ++a) out[1 + a] = beta[a];
    return MCLCAR_OK;
}

}  // namespace

// sum_i log(1 - rho_k lam_i) for K values of rho: on the device when there is one and the sums are long enough to pay
// for a launch and the upload of the spectrum, on the host otherwise
int logdet_sums(mclcar_ctx* ctx, const double* rho, int K, double* out) {
    MCL_TRY(ensure_spectrum(ctx));
    const int nn = (int)ctx->eigs.size();
    if (ctx->host_only || (int64_t)K * nn < 20000) {
        for (int k = 0; k < K; ++k)
            out[k] = (double)spectral_sum(ctx, [&](double l) { return (long double)std::log1p(-rho[k] * l); });
        return MCLCAR_OK;
    }
    MCL_TRY(require_device(ctx));
    double *d_e = nullptr, *d_r = nullptr, *d_o = nullptr, *d_w = nullptr;
    int st = pool_alloc(ctx, (void**)&d_e, ctx->eigs.size() * sizeof(double));
    if (!st && !ctx->eig_w.empty()) st = pool_alloc(ctx, (void**)&d_w, ctx->eig_w.size() * sizeof(double));
    if (!st) st = pool_alloc(ctx, (void**)&d_r, (size_t)K * sizeof(double));
    if (!st) st = pool_alloc(ctx, (void**)&d_o, (size_t)K * sizeof(double));
    if (!st) {
        cudaError_t e = cudaMemcpyAsync(d_e, ctx->eigs.data(), ctx->eigs.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_r, rho, (size_t)K * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && d_w) e = cudaMemcpyAsync(d_w, ctx->eig_w.data(), ctx->eig_w.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            ProfScope prof(ctx, PROF_REDUCE, 1);
            logdet_kernel<<<K, 256, 0, ctx->stream>>>(d_e, d_w, nn, d_r, d_o);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = fail(ctx, MCLCAR_ECUDA, "logdet kernel: %s", cudaGetErrorString(e));
    }
    pool_free(ctx, d_e); pool_free(ctx, d_r); pool_free(ctx, d_o); pool_free(ctx, d_w);
    return st;
}

}  // namespace mclcar

using namespace mclcar;

extern "C" {

int mclcar_dcar_sigmabeta(mclcar_ctx* ctx, const double* rho, int K, double* out) {
    if (!ctx || !rho || !out || K < 1) return MCLCAR_EINVAL;
    if (!ctx->has_obs) return fail(ctx, MCLCAR_ESTATE, "set graph, design and observations first");
    for (int k = 0; k < K; ++k) MCL_TRY(sigmabeta_one(ctx, rho[k], out + (size_t)k * (1 + ctx->p)));
    return MCLCAR_OK;
}

int mclcar_dcar_profile_eval(mclcar_ctx* ctx, mclcar_samples* s, const double* rho, int K, const double* rho_cons,
                             int shift_mode, double* lr, double* var) {
    if (!ctx || !s || !rho || !lr || K < 1) return MCLCAR_EINVAL;
    const int P = 2 + ctx->p;
    std::vector<double> psi((size_t)K * P);
    for (int k = 0; k < K; ++k) {
        psi[(size_t)k * P] = rho[k];
        MCL_TRY(sigmabeta_one(ctx, rho[k], psi.data() + (size_t)k * P + 1));
    }
    return mclcar_dcar_eval(ctx, s, psi.data(), K, P, rho_cons, nullptr, nullptr, shift_mode, lr, var);
}

int mclcar_dcar_loglik(mclcar_ctx* ctx, const double* psi, int K, int P, const double* rho_cons, double* out) {
    if (!ctx || !psi || !out || K < 1) return MCLCAR_EINVAL;
    if (!ctx->has_obs) return fail(ctx, MCLCAR_ESTATE, "set graph, design and observations first");
    const int p = ctx->p, n = ctx->n;
    if (P != 2 && P != 2 + p) return fail(ctx, MCLCAR_EINVAL, "parameter vector of length %d; expected 2 or %d", P, 2 + p);
    std::vector<double> rho(K), ld(K);
    for (int k = 0; k < K; ++k) rho[k] = psi[(size_t)k * P];
    MCL_TRY(logdet_sums(ctx, rho.data(), K, ld.data()));
    MCL_TRY(ensure_algebra(ctx));
    const double* q = ctx->qobs.data();
    for (int k = 0; k < K; ++k) {
        const double* t = psi + (size_t)k * P;
        const double r = t[0], sg = t[1];
        if (rho_cons && !(r > rho_cons[0] && r < rho_cons[1])) { out[k] = -1e8; continue; }  // R/loglik.dCAR.R:36-40
        double rr = q[0], rwr = q[1];
        for (int a = 0; a < P - 2; ++a) {
            rr -= 2.0 * t[2 + a] * q[2 + a];
            rwr -= 2.0 * t[2 + a] * q[2 + p + a];
            for (int c = 0; c < P - 2; ++c) {
                rr += t[2 + a] * ctx->G0[(size_t)a * p + c] * t[2 + c];
                rwr += t[2 + a] * ctx->G1[(size_t)a * p + c] * t[2 + c];
            }
        }
        out[k] = -0.5 * n * (std::log(2.0 * M_PI) + std::log(sg)) + 0.5 * ld[k] - (rr - r * rwr) / (2.0 * sg);
    }
    return MCLCAR_OK;
}

int mclcar_dcar_ploglik(mclcar_ctx* ctx, const double* rho, int K, double* out) {
    if (!ctx || !rho || !out || K < 1) return MCLCAR_EINVAL;
    if (!ctx->has_obs) return fail(ctx, MCLCAR_ESTATE, "set graph, design and observations first");
    std::vector<double> ld(K), sb(1 + ctx->p);
    MCL_TRY(logdet_sums(ctx, rho, K, ld.data()));
    for (int k = 0; k < K; ++k) {
        MCL_TRY(sigmabeta_one(ctx, rho[k], sb.data()));
        out[k] = -0.5 * ctx->n * std::log(sb[0]) + 0.5 * ld[k] - 0.5 * ctx->n;
    }
    return MCLCAR_OK;
}

/* Hessian.dCAR(pars, data), R/loglik.dCAR.R:171-219: the exact Hessian at a parameter value, from the p x p design
 * constants, the observation's statistics and the spectrum -- what summary.OptimMCL / summary.rsmMCL hand to vmle.dCAR
 * (R/summary.OptimMCL.R:24-26).  As coded in the reference: the rho-beta block is -X'r/sigma (:198), the rho-rho entry
 * -1/2 sum (lam/(1 - rho lam))^2 with no data term.  out: P x P, P = 2 or 2 + p.  Works on a host-only context when
 * the eigenvalues are known. */
int mclcar_dcar_hessian_exact(mclcar_ctx* ctx, const double* pars, int P, double* out) {
    if (!ctx || !pars || !out) return MCLCAR_EINVAL;
    if (!ctx->has_obs) return fail(ctx, MCLCAR_ESTATE, "set graph, design and observations first");
    const int p = ctx->p, n = ctx->n;
    if (P != 2 && P != 2 + p) return fail(ctx, MCLCAR_EINVAL, "parameter vector of length %d; expected 2 or %d", P, 2 + p);
    MCL_TRY(ensure_spectrum(ctx));
    MCL_TRY(ensure_algebra(ctx));
    const int pb = P - 2;
    const double rho = pars[0], sg = pars[1];
    const double* beta = pars + 2;
    const double* q = ctx->qobs.data();
    // residual forms from the statistics of y: r'r, r'Wr, X'r, X'Wr
    double rr = q[0], rwr = q[1];
    std::vector<double> xr(pb), xwr(pb);
    for (int a = 0; a < pb; ++a) {
        xr[a] = q[2 + a];
        xwr[a] = q[2 + p + a];
        rr -= 2.0 * beta[a] * q[2 + a];
        rwr -= 2.0 * beta[a] * q[2 + p + a];
        for (int c = 0; c < pb; ++c) {
            const double g0 = ctx->G0[(size_t)a * p + c], g1 = ctx->G1[(size_t)a * p + c];
            xr[a] -= g0 * beta[c];
            xwr[a] -= g1 * beta[c];
            rr += beta[a] * g0 * beta[c];
            rwr += beta[a] * g1 * beta[c];
        }
    }
    for (int t = 0; t < P * P; ++t) out[t] = 0.0;
    out[0] = -0.5 * (double)spectral_sum(ctx, [&](double l) { const long double u = l / (1.0L - rho * l); return u * u; });
    out[1] = out[P] = -rwr / (2.0 * sg * sg);
    out[P + 1] = n / (2.0 * sg * sg) - (rr - rho * rwr) / (sg * sg * sg);
    for (int a = 0; a < pb; ++a) {
        out[2 + a] = out[(size_t)(2 + a) * P] = -xr[a] / sg;                                   // as coded at :198
        out[P + 2 + a] = out[(size_t)(2 + a) * P + 1] = -(xr[a] - rho * xwr[a]) / (sg * sg);
        for (int c = a; c < pb; ++c)
            out[(size_t)(2 + a) * P + 2 + c] = out[(size_t)(2 + c) * P + 2 + a] =
                -(ctx->G0[(size_t)a * p + c] - rho * ctx->G1[(size_t)a * p + c]) / sg;
    }
    return MCLCAR_OK;
}

/* The Monte Carlo likelihood surface: mcl.dCAR(c(rho_i, sigma_j, beta), ..., Evar = TRUE) on the whole nr x ns grid
 * expand.grid(r.vec, s.vec) that rsmMCL evaluates loglik.dCAR on for its plots (R/rsmMCL.R:26-28,95-102) -- rho varies
 * fastest, lr[i + nr * j], the layout of `Exacts[[i]]$lrs` that plot.RSM draws (R/rsmSub.R:315-327).  One statistics
 * pass (already done for the sample set) and one reduction launch per 16384 grid points. */
int mclcar_dcar_surface(mclcar_ctx* ctx, mclcar_samples* s, const double* rho, int nr, const double* sigma, int ns,
                        const double* beta, int pb, const double* rho_cons, int shift_mode, double* lr, double* var) {
    if (!ctx || !s || !rho || !sigma || !lr || nr < 1 || ns < 1 || pb < 0 || (pb > 0 && !beta)) return MCLCAR_EINVAL;
    const int P = 2 + pb;
    const int64_t K = (int64_t)nr * ns, chunk = 16384;
    std::vector<double> psi((size_t)std::min(K, chunk) * P);
    for (int64_t k0 = 0; k0 < K; k0 += chunk) {
        const int kc = (int)std::min(chunk, K - k0);
        for (int k = 0; k < kc; ++k) {
            double* t = psi.data() + (size_t)k * P;
            t[0] = rho[(k0 + k) % nr];
            t[1] = sigma[(k0 + k) / nr];
            for (int a = 0; a < pb; ++a) t[2 + a] = beta[a];
        }
        MCL_TRY(mclcar_dcar_eval(ctx, s, psi.data(), kc, P, rho_cons, nullptr, nullptr, shift_mode, lr + k0, var ? var + k0 : nullptr));
    }
    return MCLCAR_OK;
}

}  // extern "C"

// Hierarchical CAR (Gaussian response): mirrors R/HCAR.sim.R and R/mcl.HCAR.R.
//
//   y = X beta + Z u + e,   e ~ N(0, sigma.e (I - rho W)^-1),   u ~ N(0, sigma.u (I - lambda M)^-1)
// A sample is a district-effect vector u (K); with r0 = y - Z u and e = r0 - X beta every
// quantity the reference forms per sample (log.uy, grad.uy, Hessian.uy) is affine in
//   q = ( r0'r0, r0'W r0, X'r0 [p], X'W r0 [p], u'u, u'Mu ),
// and these reduce to K-dimensional sparse forms in u:
//   r0'r0  = y'y  - 2 (Z'y)'u  + u'(Z'Z)u        X'r0  = X'y  - (Z'X)'u
//   r0'Wr0 = y'Wy - 2 (Z'Wy)'u + u'(Z'WZ)u       X'Wr0 = X'Wy - (Z'WX)'u
// computed by the site-major quadratic-form kernel (stats_csr.cu) on the K x N sample matrix
// -- the n-vector residual of the reference (R/mcl.HCAR.R:64-69, one per sample and per
// evaluation) is never formed.  The u | y draw of sim.HCAR (R/HCAR.sim.R:64-69, one Cholesky
// of Q* = Z'Qe Z + Qu and N backsolves) is replaced by the chromatic Gibbs sampler on Q*.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "common.cuh"
#include "engine.hpp"

namespace mclcar {

struct HostCsr {
    int rows = 0, cols = 0;
    std::vector<int> rowptr, colidx;
    std::vector<double> val;
};

struct HcarModel {
    int K = 0;
    bool has_W = true;
    HostCsr M, Z, ZtZ, ZtWZ;
    std::vector<double> bb;                       // eigenvalues of M
    std::vector<double> Zty, ZtWy, ZtX, ZtWX;     // K, K, K x p, K x p (column-major)
    double yy = 0, yWy = 0;
    std::vector<double> Xty, XtWy;                // p
    // device copies
    int *dM_rp = nullptr, *dM_ci = nullptr, *dA_rp = nullptr, *dA_ci = nullptr, *dB_rp = nullptr, *dB_ci = nullptr;
    double *dM_v = nullptr, *dA_v = nullptr, *dB_v = nullptr, *d_lin = nullptr;  // d_lin: [2+2p][K]
};

namespace {


HostCsr transpose_times(const HostCsr& A, const HostCsr& B) {  // A'B, A and B with the same number of rows
    std::vector<std::map<int, double>> acc(A.cols);
    for (int i = 0; i < A.rows; ++i)
        for (int ea = A.rowptr[i]; ea < A.rowptr[i + 1]; ++ea)
            for (int eb = B.rowptr[i]; eb < B.rowptr[i + 1]; ++eb) acc[A.colidx[ea]][B.colidx[eb]] += A.val[ea] * B.val[eb];
    HostCsr C;
    C.rows = A.cols; C.cols = B.cols;
    C.rowptr.assign(C.rows + 1, 0);
    for (int k = 0; k < C.rows; ++k) {
        for (auto& kv : acc[k]) {
            C.colidx.push_back(kv.first);
            C.val.push_back(kv.second);
        }
        C.rowptr[k + 1] = (int)C.colidx.size();
    }
    return C;
}

HostCsr w_times(const mclcar_ctx* ctx, const HostCsr& Z) {  // W Z (n x K)
    std::vector<std::map<int, double>> acc(ctx->n);
    for (int i = 0; i < ctx->n; ++i)
        for (int e = ctx->rowptr[i]; e < ctx->rowptr[i + 1]; ++e) {
            const int j = ctx->colidx[e];
            const double w = ctx->binary ? 1.0 : ctx->val[e];
            for (int ez = Z.rowptr[j]; ez < Z.rowptr[j + 1]; ++ez) acc[i][Z.colidx[ez]] += w * Z.val[ez];
        }
    HostCsr C;
    C.rows = ctx->n; C.cols = Z.cols;
    C.rowptr.assign(C.rows + 1, 0);
    for (int i = 0; i < C.rows; ++i) {
        for (auto& kv : acc[i]) {
            C.colidx.push_back(kv.first);
            C.val.push_back(kv.second);
        }
        C.rowptr[i + 1] = (int)C.colidx.size();
    }
    return C;
}

void zt_vec(const HostCsr& Z, const double* v, double* out) {  // Z'v
    for (int k = 0; k < Z.cols; ++k) out[k] = 0.0;
    for (int i = 0; i < Z.rows; ++i)
        for (int e = Z.rowptr[i]; e < Z.rowptr[i + 1]; ++e) out[Z.colidx[e]] += Z.val[e] * v[i];
}

void w_vec(const mclcar_ctx* c, const double* x, double* y) {
    for (int i = 0; i < c->n; ++i) {
        double a = 0.0;
        for (int e = c->rowptr[i]; e < c->rowptr[i + 1]; ++e) a += (c->binary ? 1.0 : c->val[e]) * x[c->colidx[e]];
        y[i] = a;
    }
}

template <class T>
int up(mclcar_ctx* ctx, T*& d, const std::vector<T>& h) {
    if (d) { pool_free(ctx, d); d = nullptr; }
    if (h.empty()) return MCLCAR_OK;
    MCL_TRY(pool_alloc(ctx, (void**)&d, h.size() * sizeof(T)));
    MCL_CUDA(ctx, cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MCLCAR_OK;
}

// q rows from the raw forms: in: xx (u'u), uMu, lin[2+2p] = (Z'y, Z'Wy, Z'X_l, Z'WX_l)'u, uAu (Z'Z), uBu (Z'WZ)
__global__ void hcar_combine_kernel(const double* f1, const double* fA, const double* fB, int64_t N, int p, double yy,
                                    double yWy, const double* Xty, const double* XtWy, double* q) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double uu = f1[i], uMu = f1[N + i];
    const double zy = f1[(size_t)2 * N + i], zwy = f1[(size_t)3 * N + i];
    q[i] = yy - 2.0 * zy + fA[N + i];
    q[N + i] = yWy - 2.0 * zwy + (fB ? fB[N + i] : 0.0);
    for (int l = 0; l < p; ++l) {
        q[(size_t)(2 + l) * N + i] = Xty[l] - f1[(size_t)(4 + l) * N + i];
        q[(size_t)(2 + p + l) * N + i] = XtWy[l] - f1[(size_t)(4 + p + l) * N + i];
    }
    q[(size_t)(2 + 2 * p) * N + i] = uu;
    q[(size_t)(3 + 2 * p) * N + i] = uMu;
}

struct HcarPars {
    double rho = 0, lam = 0, se = 1, su = 1;
    const double* beta = nullptr;
};

int split_pars(mclcar_ctx* ctx, const HcarModel& hm, const double* pars, int P, HcarPars& o) {
    const int head = hm.has_W ? 4 : 3;
    if (P != head + ctx->p) return fail(ctx, MCLCAR_EINVAL, "parameter vector of length %d; expected %d", P, head + ctx->p);
    if (hm.has_W) { o.rho = pars[0]; o.lam = pars[1]; o.se = pars[2]; o.su = pars[3]; }
    else { o.rho = 0; o.lam = pars[0]; o.se = pars[1]; o.su = pars[2]; }
    o.beta = pars + head;
    if (!(o.se > 0) || !(o.su > 0)) return fail(ctx, MCLCAR_EINVAL, "sigma.e and sigma.u must be positive");
    return MCLCAR_OK;
}

// lambda inside (1/min(bb), 1/max(bb)) -- the eigenvalues of M (data$bb) -- or, without them, inside the Gershgorin bound
int check_lambda(mclcar_ctx* ctx, const HcarModel& hm, double lam) {
    if (!hm.bb.empty()) {
        const double mn = *std::min_element(hm.bb.begin(), hm.bb.end()), mx = *std::max_element(hm.bb.begin(), hm.bb.end());
        if (mn < 0 && mx > 0) {
            if (lam < 1.0 / mn || lam > 1.0 / mx) return fail(ctx, MCLCAR_ERHO, "lambda should be within %.10g %.10g", 1.0 / mn, 1.0 / mx);
            return MCLCAR_OK;
        }
    }
    double rs = 0;
    for (int k = 0; k < hm.K; ++k) {
        double a = 0;
        for (int e = hm.M.rowptr[k]; e < hm.M.rowptr[k + 1]; ++e) a += std::fabs(hm.M.val[e]);
        rs = std::max(rs, a);
    }
    if (std::fabs(lam) * rs >= 1.0)
        return fail(ctx, MCLCAR_ERHO, "lambda should be within %.10g %.10g (Gershgorin bound; supply the eigenvalues of M for the exact range)", -1.0 / rs, 1.0 / rs);
    return MCLCAR_OK;
}

// lpars = c0 + c.q ; gradient features (order of the reference's grad.uy) ; const.pars
struct HcarAffine {
    double c0 = 0, cst = 0;
    std::vector<double> c, a0, A;
};

// ld_w: sum log(1 - rho a_i) over the spectrum of W when the caller has it already (a batch forms them in one go)
void hcar_affine(const mclcar_ctx* ctx, const HcarModel& hm, const HcarPars& t, int P, bool want_grad, HcarAffine& m,
                 const double* ld_w = nullptr) {
    const int p = ctx->p, d = 4 + 2 * p, n = ctx->n, K = hm.K;
    std::vector<double> G0b(p, 0.0), G1b(p, 0.0);
    double bG0b = 0, bG1b = 0;
    for (int l = 0; l < p; ++l) {
        for (int k = 0; k < p; ++k) {
            G0b[l] += ctx->G0[(size_t)l * p + k] * t.beta[k];
            G1b[l] += ctx->G1[(size_t)l * p + k] * t.beta[k];
        }
        bG0b += t.beta[l] * G0b[l];
        bG1b += t.beta[l] * G1b[l];
    }
    const int iu = 2 + 2 * p, im = 3 + 2 * p;
    // e'e = q0 - 2 beta'a + beta'G0 beta ; e'We = q1 - 2 beta'b + beta'G1 beta
    // lpars = -( (e'e - rho e'We)/se + (u'u - lam u'Mu)/su ) / 2
    m.c.assign(d, 0.0);
    m.c0 = -(bG0b - t.rho * bG1b) / (2 * t.se);
    m.c[0] = -1.0 / (2 * t.se);
    m.c[1] = t.rho / (2 * t.se);
    for (int l = 0; l < p; ++l) {
        m.c[2 + l] = t.beta[l] / t.se;
        m.c[2 + p + l] = -t.rho * t.beta[l] / t.se;
    }
    m.c[iu] = -1.0 / (2 * t.su);
    m.c[im] = t.lam / (2 * t.su);
    // const.pars = 1/2 logdet Qe + 1/2 logdet Qu from the spectra (chol logdets at R/mcl.HCAR.R:71-75)
    long double ld = -(long double)n * std::log(t.se) - (long double)K * std::log(t.su);
    if (hm.has_W)
        ld += ld_w ? (long double)*ld_w : spectral_sum(ctx, [&](double a) { return (long double)std::log1p(-t.rho * a); });
    for (double b : hm.bb) ld += std::log1p(-t.lam * b);
    m.cst = (double)(0.5L * ld);
    if (!want_grad) return;
    m.a0.assign(P, 0.0);
    m.A.assign((size_t)P * d, 0.0);
    auto row = [&](int a) { return m.A.data() + (size_t)a * d; };
    int a = 0;
    long double sa = 0, sb = 0;
    if (hm.has_W) {
        sa += spectral_sum(ctx, [&](double x) { return (long double)(x / (1.0 - t.rho * x)); });
        // g.rho = e'We/(2 se) - sum(aa/(1 - rho aa))/2
        m.a0[a] = bG1b / (2 * t.se) - (double)sa / 2;
        row(a)[1] = 1.0 / (2 * t.se);
        for (int l = 0; l < p; ++l) row(a)[2 + p + l] = -t.beta[l] / t.se;
        ++a;
    }
    for (double x : hm.bb) sb += x / (1.0 - t.lam * x);
    // g.lambda = u'Mu/(2 su) - sum(bb/(1 - lam bb))/2
    m.a0[a] = -(double)sb / 2;
    row(a)[im] = 1.0 / (2 * t.su);
    ++a;
    // g.sigma.e = (e'e - rho e'We)/(2 se^2) - n/(2 se)
    m.a0[a] = (bG0b - t.rho * bG1b) / (2 * t.se * t.se) - n / (2 * t.se);
    row(a)[0] = 1.0 / (2 * t.se * t.se);
    row(a)[1] = -t.rho / (2 * t.se * t.se);
    for (int l = 0; l < p; ++l) {
        row(a)[2 + l] = -t.beta[l] / (t.se * t.se);
        row(a)[2 + p + l] = t.rho * t.beta[l] / (t.se * t.se);
    }
    ++a;
    // g.sigma.u = (u'u - lam u'Mu)/(2 su^2) - K/(2 su)
    m.a0[a] = -K / (2 * t.su);
    row(a)[iu] = 1.0 / (2 * t.su * t.su);
    row(a)[im] = -t.lam / (2 * t.su * t.su);
    ++a;
    // g.beta = X'Qe e = (X'e - rho X'We)/se
    for (int l = 0; l < p; ++l, ++a) {
        m.a0[a] = (-G0b[l] + t.rho * G1b[l]) / t.se;
        row(a)[2 + l] = 1.0 / t.se;
        row(a)[2 + p + l] = -t.rho / t.se;
    }
}

HcarModel* model_of(mclcar_ctx* ctx) { return static_cast<HcarModel*>(ctx->hcar); }

int hcar_check(mclcar_ctx* ctx, mclcar_samples* s, HcarModel** hm) {
    if (!ctx || !s) return MCLCAR_EINVAL;
    *hm = model_of(ctx);
    if (!*hm) return fail(ctx, MCLCAR_ESTATE, "call mclcar_hcar_set_model first");
    if (s->ctx != ctx) return fail(ctx, MCLCAR_EINVAL, "sample set belongs to another context");
    if (s->sites != (*hm)->K) return fail(ctx, MCLCAR_EINVAL, "sample matrix has %lld rows, the model has %d districts", (long long)s->sites, (*hm)->K);
    if ((*hm)->has_W) MCL_TRY(ensure_spectrum(ctx));
    MCL_TRY(ensure_algebra(ctx));
    return MCLCAR_OK;
}

// statistics of all samples -> s->d_q [4+2p][N]
int hcar_stats(mclcar_ctx* ctx, HcarModel& hm, mclcar_samples* s) {
    const int p = ctx->p, d = 4 + 2 * p, K = hm.K;
    if (s->stats_ready && s->d == d) return MCLCAR_OK;
    if (!s->d_M) return fail(ctx, MCLCAR_ESTATE, "sample set has no matrix");
    MCL_TRY(samples_alloc_stats(ctx, s, d));
    const int64_t N = s->N;
    const void* U = s->d_M;
    int64_t ld = s->ld;
    void* tr = nullptr;
    const size_t es = s->dtype == MCLCAR_F32 ? 4 : 8;
    if (s->layout == MCLCAR_COL_IS_SAMPLE) {  // u.y as in R (K x N, column = sample): make it site-major
        MCL_TRY(pool_alloc(ctx, &tr, (size_t)K * N * es));
        MCL_TRY(launch_transpose(ctx, s->d_M, s->dtype, N, K, s->ld, tr, N));
        U = tr; ld = N;
    }
    const int L = 2 + 2 * p;
    double *f1 = nullptr, *fA = nullptr, *fB = nullptr, *d_c = nullptr;
    int st = pool_alloc(ctx, (void**)&f1, (size_t)(2 + L) * N * sizeof(double));
    if (!st) st = pool_alloc(ctx, (void**)&fA, (size_t)2 * N * sizeof(double));
    if (!st && hm.has_W) st = pool_alloc(ctx, (void**)&fB, (size_t)2 * N * sizeof(double));
    if (!st) st = pool_alloc(ctx, (void**)&d_c, (size_t)(2 * p + 1) * sizeof(double));
    if (!st) {
        std::vector<const double*> vecs(L);
        for (int l = 0; l < L; ++l) vecs[l] = hm.d_lin + (size_t)l * K;
        st = launch_quadforms_site_major(ctx, U, s->dtype, N, ld, K, hm.M.rowptr, hm.M.colidx, hm.M.val, L, vecs.data(), f1, N);
    }
    if (!st) st = launch_quadforms_site_major(ctx, U, s->dtype, N, ld, K, hm.ZtZ.rowptr, hm.ZtZ.colidx, hm.ZtZ.val, 0, nullptr, fA, N);
    if (!st && hm.has_W) st = launch_quadforms_site_major(ctx, U, s->dtype, N, ld, K, hm.ZtWZ.rowptr, hm.ZtWZ.colidx, hm.ZtWZ.val, 0, nullptr, fB, N);
    if (!st) {
        std::vector<double> cst(2 * p + 1, 0.0);
        for (int l = 0; l < p; ++l) { cst[l] = hm.Xty[l]; cst[p + l] = hm.XtWy[l]; }
        cudaError_t e = cudaMemcpyAsync(d_c, cst.data(), cst.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = fail(ctx, MCLCAR_ECUDA, "upload failed: %s", cudaGetErrorString(e));
    }
    if (!st) {
        ProfScope prof(ctx, PROF_STATS, 1);
        hcar_combine_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(f1, fA, fB, N, p, hm.yy, hm.yWy, d_c, d_c + p, s->d_q);
        if (cudaGetLastError() != cudaSuccess) st = fail(ctx, MCLCAR_ECUDA, "hcar_combine_kernel launch failed");
    }
    pool_free(ctx, f1); pool_free(ctx, fA); pool_free(ctx, fB); pool_free(ctx, d_c); pool_free(ctx, tr);
    MCL_TRY(st);
    MCL_TRY(samples_update_qbar(ctx, s));
    s->stats_ready = true;
    return MCLCAR_OK;
}

}  // namespace

void hcar_forget(mclcar_ctx* ctx) {
    HcarModel* hm = static_cast<HcarModel*>(ctx->hcar);
    if (!hm) return;
    edge_cache_clear(ctx);   // edge lists of M, Z'Z, Z'WZ are keyed by this model's arrays
    if (!ctx->host_only) {
        void* dev[] = {hm->dM_rp, hm->dM_ci, hm->dA_rp, hm->dA_ci, hm->dB_rp, hm->dB_ci, hm->dM_v, hm->dA_v, hm->dB_v, hm->d_lin};
        for (void* d : dev) pool_free(ctx, d);
    }
    delete hm;
    ctx->hcar = nullptr;
}

}  // namespace mclcar

using namespace mclcar;

extern "C" {

int mclcar_hcar_set_model(mclcar_ctx* ctx, int K, const int* M_rowptr, const int* M_colidx, const double* M_val,
                          const int* Z_rowptr, const int* Z_colidx, const double* Z_val, const double* eig_M,
                          int has_W) {
    if (!ctx || !M_rowptr || !M_colidx || !Z_rowptr || !Z_colidx || !eig_M) return MCLCAR_EINVAL;
    if (ctx->kind == GRAPH_NONE || !ctx->has_obs) return fail(ctx, MCLCAR_ESTATE, "set graph (W, possibly empty), design and observations first");
    if (K < 1) return fail(ctx, MCLCAR_EINVAL, "K must be positive");
    MCL_TRY(ensure_algebra(ctx));
    const int n = ctx->n, p = ctx->p;
    HcarModel hm;
    hm.K = K; hm.has_W = has_W != 0;
    auto load = [&](HostCsr& A, int rows, int cols, const int* rp, const int* ci, const double* v) -> int {
        A.rows = rows; A.cols = cols;
        if (rp[0] != 0) return fail(ctx, MCLCAR_EINVAL, "rowptr[0] must be 0");
        A.rowptr.assign(rp, rp + rows + 1);
        const int nnz = rp[rows];
        A.colidx.assign(ci, ci + nnz);
        A.val.resize(nnz);
        for (int e = 0; e < nnz; ++e) {
            if (ci[e] < 0 || ci[e] >= cols) return fail(ctx, MCLCAR_EINVAL, "column index out of range");
            A.val[e] = v ? v[e] : 1.0;
        }
        return MCLCAR_OK;
    };
    MCL_TRY(load(hm.M, K, K, M_rowptr, M_colidx, M_val));
    MCL_TRY(load(hm.Z, n, K, Z_rowptr, Z_colidx, Z_val));
    hm.bb.assign(eig_M, eig_M + K);
    hm.ZtZ = transpose_times(hm.Z, hm.Z);
    std::vector<double> Wy(n), tmp(n);
    w_vec(ctx, ctx->y.data(), Wy.data());
    hm.yy = ctx->qobs[0]; hm.yWy = ctx->qobs[1];
    hm.Xty.assign(ctx->qobs.begin() + 2, ctx->qobs.begin() + 2 + p);
    hm.XtWy.assign(ctx->qobs.begin() + 2 + p, ctx->qobs.begin() + 2 + 2 * p);
    hm.Zty.resize(K); hm.ZtWy.assign(K, 0.0); hm.ZtX.resize((size_t)K * p); hm.ZtWX.assign((size_t)K * p, 0.0);
    zt_vec(hm.Z, ctx->y.data(), hm.Zty.data());
    zt_vec(hm.Z, Wy.data(), hm.ZtWy.data());
    for (int l = 0; l < p; ++l) {
        zt_vec(hm.Z, ctx->X.data() + (size_t)l * n, hm.ZtX.data() + (size_t)l * K);
        zt_vec(hm.Z, ctx->WX.data() + (size_t)l * n, hm.ZtWX.data() + (size_t)l * K);
    }
    if (hm.has_W) {
        HostCsr WZ = w_times(ctx, hm.Z);
        hm.ZtWZ = transpose_times(hm.Z, WZ);
    }
    if (!ctx->host_only) {
        MCL_TRY(require_device(ctx));
        MCL_TRY(up(ctx, hm.dM_rp, hm.M.rowptr)); MCL_TRY(up(ctx, hm.dM_ci, hm.M.colidx)); MCL_TRY(up(ctx, hm.dM_v, hm.M.val));
        MCL_TRY(up(ctx, hm.dA_rp, hm.ZtZ.rowptr)); MCL_TRY(up(ctx, hm.dA_ci, hm.ZtZ.colidx)); MCL_TRY(up(ctx, hm.dA_v, hm.ZtZ.val));
        if (hm.has_W) {
            MCL_TRY(up(ctx, hm.dB_rp, hm.ZtWZ.rowptr)); MCL_TRY(up(ctx, hm.dB_ci, hm.ZtWZ.colidx)); MCL_TRY(up(ctx, hm.dB_v, hm.ZtWZ.val));
        }
        std::vector<double> lin((size_t)(2 + 2 * p) * K);
        std::copy(hm.Zty.begin(), hm.Zty.end(), lin.begin());
        std::copy(hm.ZtWy.begin(), hm.ZtWy.end(), lin.begin() + K);
        std::copy(hm.ZtX.begin(), hm.ZtX.end(), lin.begin() + (size_t)2 * K);
        std::copy(hm.ZtWX.begin(), hm.ZtWX.end(), lin.begin() + (size_t)(2 + p) * K);
        MCL_TRY(up(ctx, hm.d_lin, lin));
    }
    hcar_forget(ctx);
    ctx->hcar = new HcarModel(std::move(hm));
    return MCLCAR_OK;
}

int mclcar_hcar_attach(mclcar_ctx* ctx, mclcar_samples* s, const double* psi0, int P, const double* lpsi) {
    HcarModel* hm = nullptr;
    MCL_TRY(hcar_check(ctx, s, &hm));
    MCL_TRY(require_device(ctx));
    if (!psi0) return MCLCAR_EINVAL;
    HcarPars t;
    MCL_TRY(split_pars(ctx, *hm, psi0, P, t));
    s->psi0.assign(psi0, psi0 + P);
    if (s->d_base) { pool_free(ctx, s->d_base); s->d_base = nullptr; }
    s->base_mean = 0.0;
    if (lpsi) {
        MCL_TRY(pool_alloc(ctx, (void**)&s->d_base, (size_t)s->N * sizeof(double)));
        MCL_CUDA(ctx, cudaMemcpyAsync(s->d_base, lpsi, (size_t)s->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        long double a = 0;
        for (int64_t i = 0; i < s->N; ++i) a += lpsi[i];
        s->base_mean = (double)(a / s->N);
    }
    s->family = MCLCAR_FAMILY_HCAR;
    s->attached = true;
    return MCLCAR_OK;
}

/* the 4 + 2p sufficient statistics of every sample, q_out N x (4 + 2p) column-major (may be NULL: just form them):
 * (r0'r0, r0'W r0, X'r0[p], X'W r0[p], u'u, u'Mu) with r0 = y - Zu */
int mclcar_hcar_stats(mclcar_ctx* ctx, mclcar_samples* s, double* q_out) {
    HcarModel* hm = nullptr;
    MCL_TRY(hcar_check(ctx, s, &hm));
    MCL_TRY(require_device(ctx));
    MCL_TRY(hcar_stats(ctx, *hm, s));
    if (q_out) {
        MCL_CUDA(ctx, cudaMemcpyAsync(q_out, s->d_q, (size_t)s->d * s->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MCLCAR_OK;
}

/* lpsi_i and const.psi of sim.HCAR's return list (R/HCAR.sim.R:71-85) */
int mclcar_hcar_lpsi(mclcar_ctx* ctx, mclcar_samples* s, double* lpsi, double* const_psi) {
    HcarModel* hm = nullptr;
    MCL_TRY(hcar_check(ctx, s, &hm));
    MCL_TRY(require_device(ctx));
    if (!s->attached) return fail(ctx, MCLCAR_ESTATE, "attach psi0 first");
    HcarPars t;
    const int P = (int)s->psi0.size();
    MCL_TRY(split_pars(ctx, *hm, s->psi0.data(), P, t));
    HcarAffine m;
    hcar_affine(ctx, *hm, t, P, false, m);
    if (const_psi) *const_psi = m.cst;
    if (lpsi) {
        if (s->d_base) {
            MCL_CUDA(ctx, cudaMemcpyAsync(lpsi, s->d_base, (size_t)s->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        } else {
            MCL_TRY(hcar_stats(ctx, *hm, s));
            std::vector<double> q((size_t)s->d * s->N);
            MCL_CUDA(ctx, cudaMemcpyAsync(q.data(), s->d_q, q.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            for (int64_t i = 0; i < s->N; ++i) {
                double h = m.c0;
                for (int j = 0; j < s->d; ++j) h += m.c[j] * q[(size_t)j * s->N + i];
                lpsi[i] = h;
            }
        }
    }
    return MCLCAR_OK;
}

static int hcar_run(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int Kc, int P, int what, int evar,
                    int shift_mode, double* out0, double* out1, int* n_clamped) {
    HcarModel* hmp = nullptr;
    MCL_TRY(hcar_check(ctx, s, &hmp));
    HcarModel& hm = *hmp;
    MCL_TRY(require_device(ctx));
    if (!psi || Kc < 1 || !out0) return MCLCAR_EINVAL;
    if (!s->attached) return fail(ctx, MCLCAR_ESTATE, "call mclcar_hcar_attach (or _prep) first");
    if (P > MCLCAR_MAX_P) return fail(ctx, MCLCAR_ELIMIT, "P = %d exceeds MCLCAR_MAX_P", P);
    MCL_TRY(hcar_stats(ctx, hm, s));
    const int p = ctx->p, d = s->d, n = ctx->n, K = hm.K;
    const int F = what == WHAT_LR ? 0 : P;
    HcarPars t0, t;
    MCL_TRY(split_pars(ctx, hm, s->psi0.data(), (int)s->psi0.size(), t0));
    // log det(I - rho W) of the importance point and of every candidate in one go: n terms each -- with n = 10^4
    // individuals the host sums were 1.5 of the 1.8 ms of a 16-candidate batch
    std::vector<double> rho_all(Kc + 1), ldw_all(Kc + 1);
    rho_all[0] = t0.rho;
    for (int k = 0; k < Kc; ++k) {
        MCL_TRY(split_pars(ctx, hm, psi + (size_t)k * P, P, t));
        rho_all[1 + k] = t.rho;
    }
    if (hm.has_W) MCL_TRY(logdet_sums(ctx, rho_all.data(), Kc + 1, ldw_all.data()));
    HcarAffine m0, m;
    hcar_affine(ctx, hm, t0, P, false, m0, hm.has_W ? &ldw_all[0] : nullptr);
    ReducePlan plan;
    plan.d_q = s->d_q; plan.ldq = s->N; plan.d = d; plan.N = s->N; plan.K = Kc; plan.F = F; plan.what = what;
    // the reference takes exp() without any shift (R/mcl.HCAR.R:78); the max shift tells us
    // whether that exp would have overflowed / underflowed (clamp at :81-84)
    plan.shift_mode = what == WHAT_LR ? MCLCAR_SHIFT_MAX : shift_mode;
    plan.d_base = s->d_base;
    const int cs = coef_stride_of(d, F);
    plan.coef_stride = cs;
    plan.coef.assign((size_t)Kc * cs, 0.0);
    std::vector<double> cst(Kc);
    for (int k = 0; k < Kc; ++k) {
        MCL_TRY(split_pars(ctx, hm, psi + (size_t)k * P, P, t));
        hcar_affine(ctx, hm, t, P, what != WHAT_LR, m, hm.has_W ? &ldw_all[1 + k] : nullptr);
        cst[k] = m.cst;
        double* cf = plan.coef.data() + (size_t)k * cs;
        cf[0] = m.c0;
        for (int j = 0; j < d; ++j) cf[1 + j] = m.c[j];
        if (!s->d_base) {
            cf[0] -= m0.c0;
            for (int j = 0; j < d; ++j) cf[1 + j] -= m0.c[j];
        }
        double shift = cf[0];
        for (int j = 0; j < d; ++j) shift += cf[1 + j] * s->qbar[j];
        if (s->d_base) shift -= s->base_mean;
        cf[coef_off_shift(d)] = shift;
        cf[coef_off_n(d)] = (double)s->N;
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
    const int64_t T = tuple_len(what, F, d);
    std::vector<double> tuples((size_t)Kc * T);
    MCL_TRY(run_reduce(ctx, plan, tuples.data(), true));
    MCL_TRY(gather_and_merge(ctx, what, F, d, Kc, tuples.data()));
    if (n_clamped) *n_clamped = 0;
    std::vector<double> vbar(1 + F), C((size_t)(1 + F) * (1 + F));
    for (int k = 0; k < Kc; ++k) {
        const double* tup = tuples.data() + (size_t)k * T;
        const double nn = tup[0], shift = tup[1];
        MCL_TRY(split_pars(ctx, hm, psi + (size_t)k * P, P, t));
        if (what == WHAT_LR) {
            centred_moments(tup, 1, vbar.data(), C.data());
            const double dconst = cst[k] - m0.cst;  // const.pars - const.psi
            double lr = dconst + shift + std::log(vbar[0]);
            // exp(lpars + const.pars - lpsi - const.psi) of the largest term: Inf / all zero in the reference
            const double lmax = dconst + shift;  // max shift: the largest log-ratio
            if (lmax > 709.782712893384) { lr = 1e5; if (n_clamped) ++*n_clamped; }
            else if (lmax < -745.1332191019412) { lr = -1e5; if (n_clamped) ++*n_clamped; }
            out0[k] = lr;
            if (out1) out1[k] = C[0] / (nn - 1.0) / (vbar[0] * vbar[0]) / nn;
        } else if (what == WHAT_GRAD) {
            const int V = 1 + P;
            centred_moments(tup, V, vbar.data(), C.data());
            const double eb = vbar[0];
            for (int a = 0; a < P; ++a) out0[(size_t)k * P + a] = vbar[1 + a] / eb;
            if (out1 && evar)
                for (int a = 0; a < P; ++a)
                    for (int b = 0; b < P; ++b)
                        out1[(size_t)k * P * P + a * P + b] = C[(size_t)(1 + a) * V + 1 + b] / (nn - 1.0) / (eb * eb) / nn;
        } else {  // Hessian, R/mcl.HCAR.R:347-388
            const double Se = tup[2];
            const double* Sf = tup + 3;
            const double* Sff = tup + 3 + P;
            std::vector<double> gw(P);
            for (int a = 0; a < P; ++a) gw[a] = Sf[a] / Se;
            double* o = out0 + (size_t)k * P * P;
            for (int x = 0; x < P * P; ++x) o[x] = 0.0;
            long double sa2 = 0, sb2 = 0, sa1 = 0, sb1 = 0;
            for (double x : hm.bb) { const double r = x / (1.0 - t.lam * x); sb1 += r; sb2 += r * r; }
            if (hm.has_W) {
                sa1 = spectral_sum(ctx, [&](double x) { return (long double)(x / (1.0 - t.rho * x)); });
                sa2 = spectral_sum(ctx, [&](double x) { const double r = x / (1.0 - t.rho * x); return (long double)(r * r); });
            }
            const int head = hm.has_W ? 4 : 3;
            const int il = hm.has_W ? 1 : 0, ie = il + 1, iuu = il + 2;
            // weighted means of the raw forms recovered from the gradient features
            const double uMu = 2 * t.su * (gw[il] + (double)sb1 / 2);
            const double uQu = 2 * t.su * (gw[iuu] + K / (2 * t.su));               // u'Qu u
            const double eQe = 2 * t.se * (gw[ie] + n / (2 * t.se));                 // e'Qe e (W NULL: e'e/se)
            o[il * P + il] = -(double)sb2 / 2;
            o[il * P + iuu] = o[iuu * P + il] = -uMu / (2 * t.su * t.su);
            o[iuu * P + iuu] = -uQu / (t.su * t.su) + K / (2 * t.su * t.su);
            o[ie * P + ie] = -eQe / (t.se * t.se) + n / (2 * t.se * t.se);
            if (hm.has_W) {
                const double eWe = 2 * t.se * (gw[0] + (double)sa1 / 2);
                o[0] = -(double)sa2 / 2;
                o[0 * P + ie] = o[ie * P + 0] = -eWe / (2 * t.se * t.se);
                for (int l = 0; l < p; ++l) {
                    o[0 * P + head + l] = o[(head + l) * P + 0] = -gw[head + l];  // H.rho.beta = -X'Qe e as coded (:356)
                    o[ie * P + head + l] = o[(head + l) * P + ie] = -gw[head + l] / t.se;
                }
            } else {
                for (int l = 0; l < p; ++l)  // H.sig.beta = -X'e/se^2 = -g.beta/se
                    o[ie * P + head + l] = o[(head + l) * P + ie] = -gw[head + l] / t.se;
            }
            for (int l = 0; l < p; ++l)
                for (int r = 0; r < p; ++r)
                    o[(head + l) * P + head + r] = -(ctx->G0[(size_t)l * p + r] - t.rho * ctx->G1[(size_t)l * p + r]) / t.se;
            int sidx = 0;
            for (int a = 0; a < P; ++a)
                for (int b = a; b < P; ++b, ++sidx) {
                    const double v = Sff[sidx] / Se - gw[a] * gw[b];
                    o[a * P + b] += v;
                    if (a != b) o[b * P + a] += v;
                }
        }
    }
    return MCLCAR_OK;
}

int mclcar_hcar_eval(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, double* lr, double* var,
                     int* n_clamped) {
    return hcar_run(ctx, s, psi, K, P, WHAT_LR, 0, MCLCAR_SHIFT_MAX, lr, var, n_clamped);
}
int mclcar_hcar_grad(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int evar, int shift_mode,
                     double* grad, double* gradvar) {
    return hcar_run(ctx, s, psi, K, P, WHAT_GRAD, evar, shift_mode, grad, gradvar, nullptr);
}
int mclcar_hcar_hess(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int shift_mode, double* hess) {
    return hcar_run(ctx, s, psi, K, P, WHAT_HESS, 0, shift_mode, hess, nullptr, nullptr);
}

/* sim.HCAR(psi, data, n.samples), R/HCAR.sim.R:3-87 */
int mclcar_hcar_prep(mclcar_ctx* ctx, const double* psi0, int P, int64_t N, const mclcar_sampler_opts* opts,
                     mclcar_samples** out) {
    if (!ctx || !psi0 || !out) return MCLCAR_EINVAL;
    *out = nullptr;
    MCL_TRY(require_device(ctx));
    HcarModel* hmp = model_of(ctx);
    if (!hmp) return fail(ctx, MCLCAR_ESTATE, "call mclcar_hcar_set_model first");
    HcarModel& hm = *hmp;
    if (hm.has_W) MCL_TRY(ensure_spectrum(ctx));
    HcarPars t;
    MCL_TRY(split_pars(ctx, hm, psi0, P, t));
    if (N < 1 || N >= ((int64_t)1 << 31)) return fail(ctx, MCLCAR_EINVAL, "bad sample count");
    // Q* = Z'Qe Z + Qu is positive definite when Qe and Qu are: the reference's chol() of either stops otherwise
    // (R/HCAR.sim.R:37-40,66) -- a positive diagonal of Q* alone would let an indefinite Q* through and the chain diverge
    if (hm.has_W) MCL_TRY(check_rho(ctx, t.rho));
    MCL_TRY(check_lambda(ctx, hm, t.lam));
    const int K = hm.K, p = ctx->p;
    // Q* = (Z'Z - rho Z'WZ)/se + (I - lam M)/su ;  b = ((Z'y - rho Z'Wy) - (Z'X - rho Z'WX) beta)/se
    std::vector<std::map<int, double>> rows(K);
    for (int k = 0; k < K; ++k) {
        rows[k][k] += 1.0 / t.su;
        for (int e = hm.ZtZ.rowptr[k]; e < hm.ZtZ.rowptr[k + 1]; ++e) rows[k][hm.ZtZ.colidx[e]] += hm.ZtZ.val[e] / t.se;
        if (hm.has_W)
            for (int e = hm.ZtWZ.rowptr[k]; e < hm.ZtWZ.rowptr[k + 1]; ++e) rows[k][hm.ZtWZ.colidx[e]] -= t.rho * hm.ZtWZ.val[e] / t.se;
        for (int e = hm.M.rowptr[k]; e < hm.M.rowptr[k + 1]; ++e) rows[k][hm.M.colidx[e]] -= t.lam * hm.M.val[e] / t.su;
    }
    std::vector<int> rp(K + 1, 0), ci;
    GmrfHost g;
    g.n = K;
    g.qdiag.assign(K, 0.0);
    g.b.assign(K, 0.0);
    double jac = 0.0;  // Gershgorin bound on the Jacobi contraction factor
    for (int k = 0; k < K; ++k) {
        double off = 0.0;
        for (auto& kv : rows[k]) {
            if (kv.first == k) { g.qdiag[k] = kv.second; continue; }
            ci.push_back(kv.first);
            g.qoff.push_back(kv.second);
            off += std::fabs(kv.second);
        }
        rp[k + 1] = (int)ci.size();
        if (!(g.qdiag[k] > 0)) return fail(ctx, MCLCAR_ERHO, "the posterior precision Z'Qe Z + Qu is not positive definite at district %d", k);
        jac = std::max(jac, off / g.qdiag[k]);
        double bk = hm.Zty[k] - t.rho * hm.ZtWy[k];
        for (int l = 0; l < p; ++l) bk -= (hm.ZtX[(size_t)l * K + k] - t.rho * hm.ZtWX[(size_t)l * K + k]) * t.beta[l];
        g.b[k] = bk / t.se;
    }
    // greedy colouring of the Q* pattern
    std::vector<int> color(K, -1), mark(K + 1, -1);
    int nc = 0;
    for (int k = 0; k < K; ++k) {
        for (int e = rp[k]; e < rp[k + 1]; ++e)
            if (color[ci[e]] >= 0) mark[color[ci[e]]] = k;
        int c = 0;
        while (mark[c] == k) ++c;
        color[k] = c;
        nc = std::max(nc, c + 1);
    }
    std::vector<int> cptr(nc + 1, 0), order(K);
    for (int k = 0; k < K; ++k) cptr[color[k] + 1]++;
    for (int c = 0; c < nc; ++c) cptr[c + 1] += cptr[c];
    {
        std::vector<int> pos(cptr.begin(), cptr.end() - 1);
        for (int k = 0; k < K; ++k) order[pos[color[k]]++] = k;
    }
    g.rowptr = &rp; g.colidx = &ci; g.color_ptr = &cptr; g.order = &order;

    mclcar_sampler_opts o{};
    if (opts) o = *opts;
    o.keep = 1;
    if (o.dtype != MCLCAR_F64) o.dtype = MCLCAR_F32;
    int burn = o.burn, thin = o.thin;
    const double gam = std::min(std::max(jac * jac, 1e-3), 0.9995);
    if (thin <= 0) thin = (int)std::min(4000.0, std::max(1.0, std::ceil(std::log(0.01) / std::log(gam))));
    if (burn <= 0) burn = (int)std::min(40000.0, std::max(8.0, std::ceil(std::log(1e-6) / std::log(gam))));
    const size_t es = o.dtype == MCLCAR_F32 ? 4 : 8;
    mclcar_samples* s = new (std::nothrow) mclcar_samples();
    if (!s) return fail(ctx, MCLCAR_ENOMEM, "out of host memory");
    s->ctx = ctx; s->dtype = o.dtype; s->sites = K; s->N = N; s->owned = true; s->layout = MCLCAR_ROW_IS_SAMPLE;
    const int64_t align = 16 / es;
    s->ld = (N + align - 1) / align * align;
    int st = pool_alloc(ctx, &s->d_M, (size_t)K * s->ld * es);
    const int64_t C = o.n_chains > 0 ? o.n_chains + (o.n_chains & 1) : default_csr_chains(ctx, K, N, burn, thin);
    const double omega = o.omega > 0 ? o.omega : 1.0;
    if (st == MCLCAR_OK)
        st = run_csr_sampler(ctx, g, FAM_GAUSS, omega, nullptr, nullptr, nullptr, nullptr, N, C, burn, thin, s->d_M,
                             o.dtype == MCLCAR_F64, s->ld, &s->accept, &s->sweeps);
    s->chains = C; s->burn = burn; s->thin = thin; s->omega = omega;
    if (st == MCLCAR_OK) st = mclcar_hcar_attach(ctx, s, psi0, P, nullptr);
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

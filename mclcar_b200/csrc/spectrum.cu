// Spectral sums of a CSR neighbourhood matrix without its dense spectrum.
//
// The reference needs three scalars of W per candidate rho: 1/2 log det(I - rho W) (chol.spam, R/mcl.bin.R:22-24,50-52;
// R/mcl.HCAR.R:71-75), sum lam/(1 - rho lam) and sum (lam/(1 - rho lam))^2 (eigen(W), R/mcl.bin.R:176,192,265;
// R/mcl.HCAR.R:166-173) -- a dense O(n^3) decomposition per CALL, impossible beyond a few thousand sites.  All three
// are traces tr f(W).  Stochastic Lanczos quadrature: for m Rademacher probes v, k Lanczos steps on W give a
// tridiagonal T_v whose eigenvalues theta and squared first eigenvector components tau^2 are a Gauss quadrature of
// v'f(W)v; tr f(W) ~ (n/m) sum_v sum_i tau_vi^2 f(theta_vi).  The Lanczos runs do NOT depend on rho or f: they run
// ONCE per graph, all probes at once (probe = lane, the site-major layout of the sample matrices; one SpMV kernel per
// step over the context's CSR), and leave m k weighted nodes that serve every later candidate exactly like a spectrum.
// The weights are then calibrated so that the moments the graph gives exactly -- tr I = n, tr W = sum of the
// diagonal, tr W^2 = ||W||_F^2 -- are reproduced exactly (a quadratic reweighting): the estimator's error is left
// with the third-order remainder of f.  Stated tolerance (tests/test_gpu_spectrum.py, n = 2000 and 5000 against dense
// eigvalsh): 2e-3 relative on log det(I - rho W), 1 % of the CHANGE of the log-determinant between two candidates
// (what the likelihood ratio sees), 2e-2 relative on the two gradient / Hessian sums, extreme eigenvalues to 1e-4.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "engine.hpp"

namespace mclcar {
namespace {

constexpr int kM = 64;   // probes (columns); fixed: the kernels below assume 64 columns per row

// y[s][j] = sum_e W_se x[col_e][j]
__global__ void __launch_bounds__(256) slq_spmv_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                       const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * kM) return;
    const int s = (int)(idx / kM), j = (int)(idx % kM);
    double a = 0.0;
    for (int e = rowptr[s]; e < rowptr[s + 1]; ++e) a = fma(val ? val[e] : 1.0, x[(size_t)colidx[e] * kM + j], a);
    y[idx] = a;
}

// column dot products over a chunk of rows: partial[block][j] = sum_{rows of the block} a[r][j] b[r][j]  (fixed order)
__global__ void __launch_bounds__(256) slq_coldot_kernel(int n, int rows_per_block, const double* __restrict__ a,
                                                         const double* __restrict__ b, double* __restrict__ partial) {
    __shared__ double sh[4][kM];
    const int j = threadIdx.x % kM, rl = threadIdx.x / kM;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(n, r0 + rows_per_block);
    double acc = 0.0;
    for (int r = r0 + rl; r < r1; r += 4) acc = fma(a[(size_t)r * kM + j], b[(size_t)r * kM + j], acc);
    sh[rl][j] = acc;
    __syncthreads();
    if (rl == 0) partial[(size_t)blockIdx.x * kM + j] = (sh[0][j] + sh[1][j]) + (sh[2][j] + sh[3][j]);
}
__global__ void slq_colsum_kernel(int blocks, const double* __restrict__ partial, double* __restrict__ out) {
    const int j = threadIdx.x;
    double a = 0.0;
    for (int b = 0; b < blocks; ++b) a += partial[(size_t)b * kM + j];
    out[j] = a;
}
// w <- w - alpha v - beta v_prev   (alpha, beta per column)
__global__ void __launch_bounds__(256) slq_axpy_kernel(int n, double* __restrict__ w, const double* __restrict__ v,
                                                       const double* __restrict__ vp, const double* __restrict__ alpha,
                                                       const double* __restrict__ beta) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * kM) return;
    const int j = (int)(idx % kM);
    w[idx] = w[idx] - alpha[j] * v[idx] - beta[j] * vp[idx];
}
// v_next <- w / sqrt(nrm2)   (0 when the Krylov space is exhausted); beta_out <- sqrt(nrm2)
__global__ void __launch_bounds__(256) slq_scale_kernel(int n, const double* __restrict__ w, const double* __restrict__ nrm2,
                                                        double* __restrict__ vnext, double* __restrict__ beta_out, double tiny) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * kM) return;
    const int j = (int)(idx % kM);
    const double b = sqrt(fmax(nrm2[j], 0.0));
    vnext[idx] = b > tiny ? w[idx] / b : 0.0;
    if (idx < kM) beta_out[j] = b > tiny ? b : 0.0;
}

// eigenvalues of the symmetric tridiagonal (d, e) and the FIRST components of its normalised eigenvectors
// (implicit QL with Wilkinson shifts; only the first row of the rotation product is carried)
void tridiag_quadrature(std::vector<double>& d, std::vector<double>& e, std::vector<double>& z0) {
    const int k = (int)d.size();
    z0.assign(k, 0.0);
    z0[0] = 1.0;
    e.push_back(0.0);
    for (int l = 0; l < k; ++l) {
        int iter = 0, mm;
        do {
            for (mm = l; mm < k - 1; ++mm) {
                const double dd = std::fabs(d[mm]) + std::fabs(d[mm + 1]);
                if (std::fabs(e[mm]) <= 2.3e-16 * dd) break;
            }
            if (mm != l) {
                if (++iter > 200) break;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[mm] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = mm - 1; i >= l; --i) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = (r = std::hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[mm] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    f = z0[i + 1];
                    z0[i + 1] = s * z0[i] + c * f;
                    z0[i] = c * z0[i] - s * f;
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[mm] = 0.0;
            }
        } while (mm != l);
    }
}

// all eigenvalues of a small dense symmetric matrix (cyclic Jacobi rotations, a: n x n row-major, destroyed)
void jacobi_eigenvalues(std::vector<double>& a, int n, std::vector<double>& ev) {
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) (i == j ? diag : off) += a[(size_t)i * n + j] * a[(size_t)i * n + j];
        if (off <= 1e-32 * (diag + off)) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[(size_t)p * n + q];
                if (apq == 0.0) continue;
                const double th = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
                const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::hypot(th, 1.0));
                const double c = 1.0 / std::hypot(t, 1.0), s = t * c;
                for (int r = 0; r < n; ++r) {      // columns p, q
                    const double arp = a[(size_t)r * n + p], arq = a[(size_t)r * n + q];
                    a[(size_t)r * n + p] = c * arp - s * arq;
                    a[(size_t)r * n + q] = s * arp + c * arq;
                }
                for (int r = 0; r < n; ++r) {      // rows p, q
                    const double apr = a[(size_t)p * n + r], aqr = a[(size_t)q * n + r];
                    a[(size_t)p * n + r] = c * apr - s * aqr;
                    a[(size_t)q * n + r] = s * apr + c * aqr;
                }
            }
    }
    ev.resize(n);
    for (int i = 0; i < n; ++i) ev[i] = a[(size_t)i * n + i];
}

}  // namespace
}  // namespace mclcar

using namespace mclcar;

/* Estimate the spectral measure of the context's W by stochastic Lanczos quadrature on the device (64 probes, `steps`
 * Lanczos steps, 0 = 80) and keep it in place of the eigenvalues: every later log-determinant / gradient / Hessian sum
 * and the rho range check use it.  Needs a CSR graph on a device context; exact eigenvalues supplied later through
 * mclcar_graph_set_eigenvalues replace it. */
extern "C" int mclcar_graph_estimate_spectrum(mclcar_ctx* ctx, int steps, uint64_t seed) {
    if (!ctx) return MCLCAR_EINVAL;
    MCL_TRY(require_device(ctx));
    if (ctx->kind == GRAPH_NONE) return fail(ctx, MCLCAR_ESTATE, "set the graph first");
    const int n = ctx->n;
    if (n <= kM) {
        // a graph no larger than the probe block: a quadrature has nothing to gain over the spectrum itself (and Lanczos
        // without re-orthogonalisation loses its exactness to ghost eigenvalues here) -- Jacobi rotations on the host
        std::vector<double> a((size_t)n * n, 0.0), ev;
        for (int i = 0; i < n; ++i)
            for (int e2 = ctx->rowptr[i]; e2 < ctx->rowptr[i + 1]; ++e2)
                a[(size_t)i * n + ctx->colidx[e2]] = ctx->binary ? 1.0 : ctx->val[e2];
        jacobi_eigenvalues(a, n, ev);
        std::sort(ev.begin(), ev.end());
        ctx->eigs = ev;
        ctx->eig_w.clear();
        ctx->ew_min = ev.front();
        ctx->ew_max = ev.back();
        return MCLCAR_OK;
    }
    int k = steps > 0 ? steps : 80;
    k = std::min(k, n);
    const size_t nm = (size_t)n * kM;
    // Rademacher probes from a splitmix64 stream
    std::vector<double> v0(nm);
    uint64_t x = seed ^ 0x9E3779B97F4A7C15ull;
    auto next = [&] { x += 0x9E3779B97F4A7C15ull; uint64_t z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
    for (size_t t = 0; t < nm; t += 64) {
        const uint64_t bits = next();
        for (size_t b = 0; b < 64 && t + b < nm; ++b) v0[t + b] = (bits >> b) & 1 ? 1.0 : -1.0;
    }
    const double inv = 1.0 / std::sqrt((double)n);
    for (double& t : v0) t *= inv;
    double *d_v = nullptr, *d_vp = nullptr, *d_w = nullptr, *d_part = nullptr, *d_sc = nullptr;
    const int rows_per_block = 256;
    const int blocks = (n + rows_per_block - 1) / rows_per_block;
    int st = MCLCAR_OK;
    std::vector<double> alpha((size_t)k * kM), beta((size_t)k * kM);
    do {
        if ((st = pool_alloc(ctx, (void**)&d_v, nm * 8))) break;
        if ((st = pool_alloc(ctx, (void**)&d_vp, nm * 8))) break;
        if ((st = pool_alloc(ctx, (void**)&d_w, nm * 8))) break;
        if ((st = pool_alloc(ctx, (void**)&d_part, (size_t)blocks * kM * 8))) break;
        if ((st = pool_alloc(ctx, (void**)&d_sc, (size_t)(2 + 2 * k) * kM * 8))) break;
        cudaStream_t s = ctx->stream;
        double* d_alpha = d_sc + 2 * kM;            // [k][kM]
        double* d_beta = d_alpha + (size_t)k * kM;  // [k][kM]: beta_i = norm after step i
        double* d_tmp = d_sc;                       // [kM] scratch, d_sc + kM: zero betas for the first step
        cudaError_t e = cudaMemcpyAsync(d_v, v0.data(), nm * 8, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_vp, 0, nm * 8, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_sc, 0, (size_t)(2 + 2 * k) * kM * 8, s);
        if (e != cudaSuccess) { st = fail(ctx, MCLCAR_ECUDA, "spectrum estimate: %s", cudaGetErrorString(e)); break; }
        const unsigned g1 = (unsigned)((nm + 255) / 256);
        ProfScope prof(ctx, PROF_SPECTRUM, (int64_t)k * 7);
        for (int i = 0; i < k; ++i) {
            slq_spmv_kernel<<<g1, 256, 0, s>>>(n, ctx->d_rowptr, ctx->d_colidx, ctx->d_val, d_v, d_w);
            slq_coldot_kernel<<<blocks, 256, 0, s>>>(n, rows_per_block, d_w, d_v, d_part);
            slq_colsum_kernel<<<1, kM, 0, s>>>(blocks, d_part, d_alpha + (size_t)i * kM);
            slq_axpy_kernel<<<g1, 256, 0, s>>>(n, d_w, d_v, d_vp, d_alpha + (size_t)i * kM, i ? d_beta + (size_t)(i - 1) * kM : d_sc + kM);
            slq_coldot_kernel<<<blocks, 256, 0, s>>>(n, rows_per_block, d_w, d_w, d_part);
            slq_colsum_kernel<<<1, kM, 0, s>>>(blocks, d_part, d_tmp);
            std::swap(d_v, d_vp);                   // v_prev <- v
            slq_scale_kernel<<<g1, 256, 0, s>>>(n, d_w, d_tmp, d_v, d_beta + (size_t)i * kM, 1e-10);
        }
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(alpha.data(), d_alpha, alpha.size() * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(beta.data(), d_beta, beta.size() * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { st = fail(ctx, MCLCAR_ECUDA, "spectrum estimate: %s", cudaGetErrorString(e)); break; }
    } while (0);
    pool_free(ctx, d_v); pool_free(ctx, d_vp); pool_free(ctx, d_w); pool_free(ctx, d_part); pool_free(ctx, d_sc);
    MCL_TRY(st);
    // Gauss quadrature of every probe
    std::vector<double> nodes, weights;
    nodes.reserve((size_t)k * kM);
    weights.reserve((size_t)k * kM);
    std::vector<double> d(k), e, z0;
    for (int j = 0; j < kM; ++j) {
        int kk = k;
        for (int i = 0; i < k - 1; ++i)
            if (beta[(size_t)i * kM + j] == 0.0) { kk = i + 1; break; }     // Krylov space exhausted
        d.assign(kk, 0.0);
        e.assign(std::max(kk - 1, 0), 0.0);
        for (int i = 0; i < kk; ++i) d[i] = alpha[(size_t)i * kM + j];
        for (int i = 0; i + 1 < kk; ++i) e[i] = beta[(size_t)i * kM + j];
        tridiag_quadrature(d, e, z0);
        for (int i = 0; i < kk; ++i) {
            nodes.push_back(d[i]);
            weights.push_back((double)n / kM * z0[i] * z0[i]);
        }
    }
    // calibrate: tr I, tr W and tr W^2 are known exactly -> quadratic reweighting w (a + b theta + c theta^2)
    long double T1 = 0, T2 = 0;
    for (int i = 0; i < n; ++i)
        for (int e2 = ctx->rowptr[i]; e2 < ctx->rowptr[i + 1]; ++e2) {
            const double w = ctx->binary ? 1.0 : ctx->val[e2];
            if (ctx->colidx[e2] == i) T1 += w;
            T2 += (long double)w * w;
        }
    long double M[5] = {0, 0, 0, 0, 0};
    for (size_t i = 0; i < nodes.size(); ++i) {
        long double p = weights[i];
        for (int r = 0; r < 5; ++r) { M[r] += p; p *= nodes[i]; }
    }
    double A[9] = {(double)M[0], (double)M[1], (double)M[2], (double)M[1], (double)M[2], (double)M[3], (double)M[2], (double)M[3], (double)M[4]};
    double rhs[3] = {(double)n, (double)T1, (double)T2}, abc[3] = {1.0, 0.0, 0.0};
    {   // 3 x 3 Gaussian elimination with pivoting; keep the raw weights if the system is degenerate
        double Mx[3][4];
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) Mx[r][c] = A[r * 3 + c]; Mx[r][3] = rhs[r]; }
        bool ok = true;
        for (int c = 0; c < 3 && ok; ++c) {
            int piv = c;
            for (int r = c + 1; r < 3; ++r) if (std::fabs(Mx[r][c]) > std::fabs(Mx[piv][c])) piv = r;
            if (std::fabs(Mx[piv][c]) < 1e-12 * std::fabs(A[0])) { ok = false; break; }
            for (int t = 0; t < 4; ++t) std::swap(Mx[c][t], Mx[piv][t]);
            for (int r = c + 1; r < 3; ++r) { const double f = Mx[r][c] / Mx[c][c]; for (int t = c; t < 4; ++t) Mx[r][t] -= f * Mx[c][t]; }
        }
        if (ok) {
            for (int r = 2; r >= 0; --r) { double sacc = Mx[r][3]; for (int t = r + 1; t < 3; ++t) sacc -= Mx[r][t] * abc[t]; abc[r] = sacc / Mx[r][r]; }
            for (size_t i = 0; i < nodes.size(); ++i) weights[i] *= abc[0] + nodes[i] * (abc[1] + nodes[i] * abc[2]);
        }
    }
    ctx->eigs = nodes;
    ctx->eig_w = weights;
    ctx->ew_min = *std::min_element(nodes.begin(), nodes.end());
    ctx->ew_max = *std::max_element(nodes.begin(), nodes.end());
    return MCLCAR_OK;
}

/* what the context holds in place of eigen(W): 0 nothing, 1 exact eigenvalues (analytic or supplied), 2 a Lanczos
 * quadrature (mclcar_graph_estimate_spectrum); *nodes = number of nodes */
extern "C" int mclcar_graph_spectrum_kind(const mclcar_ctx* ctx, int* nodes) {
    if (!ctx) return 0;
    if (nodes) *nodes = (int)ctx->eigs.size();
    return ctx->eigs.empty() ? 0 : (ctx->eig_w.empty() ? 1 : 2);
}

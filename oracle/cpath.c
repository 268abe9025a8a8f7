/* C/OpenMP twin of the oracle's sufficient-statistic path, for CPU baseline timing only.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py).  "Reference algorithm, sparse" flavour of
 * the same mathematics as R/mcl.dCAR.R:11-50,168-185 of
 * /root/reference/mclcar_0.2-0 with the O(nnz) quadratic forms the reference's dense
 * Q products reduce to, mean-shift log-mean-exp (:26-30) and unbiased weight variance (:34).
 * Parity unpinned by the reference (no golden vectors exist); checked against
 * oracle/dcar.py in tests/test_oracle.py.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int cpath_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline is meant to use the host's threads */
void cpath_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* q[i*d + j], d = 2 + 2p: (Y'Y, Y'WY, X'Y[p], (WX)'Y[p]) for N sample-major rows of Y (N x n),
 * rook torus n1 x n2; X, WX column-major n x p. */
void cpath_stats_torus(const double* Y, int64_t N, int n1, int n2, const double* X, const double* WX, int p,
                       double* q) {
    const int64_t n = (int64_t)n1 * n2;
    const int d = 2 + 2 * p;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const double* y = Y + i * n;
        double xx = 0.0, xwx = 0.0;
        for (int r = 0; r < n1; ++r) {
            const double* row = y + (int64_t)r * n2;
            const double* up = y + (int64_t)((r + n1 - 1) % n1) * n2;
            for (int c = 0; c < n2; ++c) {
                const double v = row[c];
                xx += v * v;
                xwx += v * (row[c + 1 == n2 ? 0 : c + 1] + up[c]);
            }
        }
        q[i * d + 0] = xx;
        q[i * d + 1] = 2.0 * xwx;
        for (int l = 0; l < p; ++l) {
            double a = 0.0, b = 0.0;
            const double* x = X + (int64_t)l * n;
            const double* w = WX + (int64_t)l * n;
            for (int64_t s = 0; s < n; ++s) {
                a += x[s] * y[s];
                b += w[s] * y[s];
            }
            q[i * d + 2 + l] = a;
            q[i * d + 2 + p + l] = b;
        }
    }
}

/* general CSR graph (binary or weighted), same output */
void cpath_stats_csr(const double* Y, int64_t N, int64_t n, const int* rowptr, const int* colidx, const double* val,
                     const double* X, const double* WX, int p, double* q) {
    const int d = 2 + 2 * p;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const double* y = Y + i * n;
        double xx = 0.0, xwx = 0.0;
        for (int64_t s = 0; s < n; ++s) {
            double nb = 0.0;
            for (int e = rowptr[s]; e < rowptr[s + 1]; ++e) nb += (val ? val[e] : 1.0) * y[colidx[e]];
            xx += y[s] * y[s];
            xwx += y[s] * nb;
        }
        q[i * d + 0] = xx;
        q[i * d + 1] = xwx;
        for (int l = 0; l < p; ++l) {
            double a = 0.0, b = 0.0;
            for (int64_t s = 0; s < n; ++s) {
                a += X[(int64_t)l * n + s] * y[s];
                b += WX[(int64_t)l * n + s] * y[s];
            }
            q[i * d + 2 + l] = a;
            q[i * d + 2 + p + l] = b;
        }
    }
}

/* mcl.dCAR over K candidates from the statistics: s2_i = c0_k + c_k . q_i - t20_i;
 * lr_k = s1_k - (mean + log(mean(exp(s2 - mean)))); var_k = var(w)/N (unbiased). coef: K x (1+d). */
void cpath_reduce(const double* q, const double* t20, int64_t N, int d, const double* coef, const double* s1,
                  int K, double* lr, double* var) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < K; ++k) {
        const double* c = coef + (size_t)k * (1 + d);
        double mean = 0.0;
        for (int64_t i = 0; i < N; ++i) {
            double s = c[0] - t20[i];
            for (int j = 0; j < d; ++j) s += c[1 + j] * q[i * d + j];
            mean += s;
        }
        mean /= (double)N;
        double se = 0.0;
        for (int64_t i = 0; i < N; ++i) {
            double s = c[0] - t20[i];
            for (int j = 0; j < d; ++j) s += c[1 + j] * q[i * d + j];
            se += exp(s - mean);
        }
        const double em = se / (double)N;
        double m2 = 0.0;
        for (int64_t i = 0; i < N; ++i) {
            double s = c[0] - t20[i];
            for (int j = 0; j < d; ++j) s += c[1 + j] * q[i * d + j];
            const double w = exp(s - mean) / em - 1.0;
            m2 += w * w;
        }
        lr[k] = s1[k] - (mean + log(em));
        var[k] = m2 / (double)(N - 1) / (double)N;
    }
}

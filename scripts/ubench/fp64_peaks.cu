// Micro-peaks for the fp64-bound kernels of the likelihood path (BASELINE.md section 1 asks for them):
//   DFMA  : dependent-chain-free fp64 fused multiply-adds (8 independent accumulators per thread)
//   exp   : fp64 exp() of libdevice, the cost centre of the reduction kernels (one per sample x candidate)
//   glmb  : the binomial data-term arithmetic of glm_terms_kernel per (sample, site, beta): exp(-|eta|), log1p, divide
//   glmp  : the Poisson one: exp(eta)
//   glmt  : the binomial data term as glm_terms_kernel computes it since round 2: L(|eta|) = log1p(e^-|eta|) and its
//           derivative from a piecewise degree-8 table in shared memory (9 LDS.64 + 16 DFMA + F2I / I2F per evaluation)
// Prints operations per second for the whole GPU; resident data only (no memory traffic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_peaks.cu -o fp64_peaks
#include <cstdio>
#include <cuda_runtime.h>

__device__ double g_tab[160 * 9];

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed) {
    __shared__ double stab[MODE == 4 ? 160 * 9 : 1];
    if (MODE == 4) {
        for (int t = threadIdx.x; t < 160 * 9; t += blockDim.x) stab[t] = g_tab[t];
        __syncthreads();
    }
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + 1e-3 * (threadIdx.x + 8 * j);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) a[j] = fma(a[j], 0.999999, 1e-7);
            else if (MODE == 1) a[j] = exp(a[j] * 1e-3 - 0.5) ;
            else if (MODE == 2) {
                const double eta = a[j] * 0.5 - 0.3;
                const double t = exp(-fabs(eta));
                const double l1p = log1p(t);
                const double inv = 1.0 / (1.0 + t);
                a[j] = fmax(eta, 0.0) + l1p + (eta >= 0.0 ? inv : t * inv) * 1e-3;
            } else if (MODE == 4) {
                const double eta = a[j] * 0.5 - 0.3 + 0.11 * (threadIdx.x & 31);       // lanes spread over ~14 table rows
                const double aa = fmin(fabs(eta), 39.999999);
                const int kk = (int)(aa * 4.0);
                const double t = fma(aa, 8.0, -(double)(2 * kk + 1));
                const double* c = stab + kk * 9;
                double p = c[8], dp = 0.0;
#pragma unroll
                for (int d = 7; d >= 0; --d) { dp = fma(dp, t, p); p = fma(p, t, c[d]); }
                a[j] = (fmax(eta, 0.0) + p + (eta >= 0.0 ? 1.0 + 8.0 * dp : -8.0 * dp) * 1e-3) * 0.25;
            } else a[j] = exp(a[j] * 0.25 - 1.0) * 0.5;
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char* name, int sms, int iters) {
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = sms * 8;
    k<MODE><<<grid, 256>>>(out, 10, 0.5);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k<MODE><<<grid, 256>>>(out, iters, 0.5);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
    }
    const double ops = (double)grid * 256 * iters * 8;
    printf("%-36s %8.3f ms  %.3e per second  (%.1f per clock per SM at 1.9 GHz)  err=%s\n", name, best, ops / (best * 1e-3),
           ops / (best * 1e-3) / sms / 1.9e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    run<0>("DFMA", sms, 4000);
    run<1>("exp (fp64)", sms, 400);
    run<2>("binomial data term (exp, log1p, div)", sms, 200);
    run<3>("Poisson data term (exp)", sms, 400);
    {   // any smooth table will do for the timing: a degree-1 stand-in with the right layout
        static double h[160 * 9];
        for (int k2 = 0; k2 < 160; ++k2) { h[k2 * 9] = 0.6931 / (1 + k2); h[k2 * 9 + 1] = -0.1 / (1 + k2); }
        cudaMemcpyToSymbol(g_tab, h, sizeof(h));
    }
    run<4>("binomial data term (table, degree 8 + derivative)", sms, 400);
    return 0;
}

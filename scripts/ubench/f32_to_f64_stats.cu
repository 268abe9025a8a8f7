// Microbenchmark for the next round: K2 on fp32 samples is held at 0.71-0.82 of the HBM peak by one
// F2F.F64.F32 per element (XU pipe 68 % in ncu).  Which conversion strategy sustains the HBM rate?
//   A  cvt.f64.f32 per element (what stats_torus_kernel<float> does today)
//   B  integer bit manipulation (exponent rebias + mantissa shift; ALU pipe instead of XU; zero handled, no denormals)
//   C  error-free fp32 products (p = x*y, e = fma(x, y, -p)) accumulated as fp32 pairs per 32 elements,
//      converted once per group instead of once per element
// Each variant streams a resident fp32 array once and accumulates x*x and x*x_next in fp64-equivalent precision.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 f32_to_f64_stats.cu -o f32_to_f64_stats
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double bits_to_double(float f) {
    const uint32_t u = __float_as_uint(f);
    const uint32_t e = (u >> 23) & 0xffu;
    const uint32_t hi = (u & 0x80000000u) | ((e + 896u) << 20) | ((u >> 3) & 0x000fffffu);
    const uint32_t lo = u << 29;
    return e == 0 ? 0.0 : __hiloint2double((int)hi, (int)lo);
}

template <int V>
__global__ void __launch_bounds__(256) k(const float4* __restrict__ x, size_t n4, double* out) {
    double a0 = 0.0, a1 = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = x[i];
        if (V == 0) {
            const double x0 = v.x, x1 = v.y, x2 = v.z, x3 = v.w;
            a0 = fma(x0, x0, a0); a0 = fma(x1, x1, a0); a0 = fma(x2, x2, a0); a0 = fma(x3, x3, a0);
            a1 = fma(x0, x1, a1); a1 = fma(x1, x2, a1); a1 = fma(x2, x3, a1);
        } else if (V == 1) {
            const double x0 = bits_to_double(v.x), x1 = bits_to_double(v.y), x2 = bits_to_double(v.z), x3 = bits_to_double(v.w);
            a0 = fma(x0, x0, a0); a0 = fma(x1, x1, a0); a0 = fma(x2, x2, a0); a0 = fma(x3, x3, a0);
            a1 = fma(x0, x1, a1); a1 = fma(x1, x2, a1); a1 = fma(x2, x3, a1);
        } else {
            // error-free products: hi and lo parts summed separately in fp32 over this vector, promoted once
            float ph = 0.f, pl = 0.f, qh = 0.f, ql = 0.f;
            const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float p = xs[t] * xs[t];
                pl += fmaf(xs[t], xs[t], -p);
                ph += p;
                if (t < 3) {
                    const float q = xs[t] * xs[t + 1];
                    ql += fmaf(xs[t], xs[t + 1], -q);
                    qh += q;
                }
            }
            a0 += (double)ph + (double)pl;   // NOTE: ph itself rounds; a real kernel needs two-sum here -- timing probe only
            a1 += (double)qh + (double)ql;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, a0); atomicAdd(out + 1, a1); }
}

template <int V>
void run(const char* name, const float4* x, size_t n4, double* out, int sms) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaMemset(out, 0, 16);
        cudaEventRecord(a);
        k<V><<<sms * 8, 256>>>(x, n4, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (rep) best = ms < best ? ms : best;
    }
    double h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-44s %7.3f ms  %7.1f GB/s   sums %.10e %.10e  err=%s\n", name, best, n4 * 16.0 / (best * 1e-3) / 1e9, h[0], h[1],
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const size_t n4 = (size_t)1 << 28;   // 4 GiB of floats: far larger than L2
    float4* x; double* out;
    cudaMalloc(&x, n4 * 16); cudaMalloc(&out, 16);
    cudaMemset(x, 0x3c, n4 * 16);        // 0x3c3c3c3c = 0.0115 as float
    run<0>("A cvt.f64.f32 per element", x, n4, out, p.multiProcessorCount);
    run<1>("B integer bit manipulation", x, n4, out, p.multiProcessorCount);
    run<2>("C fp32 error-free products, convert per vector", x, n4, out, p.multiProcessorCount);
    return 0;
}

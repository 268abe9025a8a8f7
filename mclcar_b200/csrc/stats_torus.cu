// K2 (torus) -- per-sample sufficient statistics in ONE pass over the sample matrix.
//
// Replaces the per-sample dense products of logunnorm / D.log.unorm
// (R/mcl.dCAR.R:168-210: n x n Q rebuilt per sample and per psi) by
//     q_i = (Y_i'Y_i,  Y_i'W Y_i,  X'Y_i [p],  (WX)'Y_i [p])
// for the rook torus of CAR.simTorus (R/CAR.simLM.R:4-7), W never materialised:
//     Y'WY = 2 * sum_sites y(r,c) * (y(r,c+1) + y(r-1,c)).
//
// Layout: sample-major ("column = sample" of the R matrix simY): sample i is the contiguous
// run M[i*ld .. i*ld + n1*n2), row-major n1 x n2.
//
// Kernel: persistent CTAs, work item = (sample, row strip).  Rows are streamed through a
// ring of shared-memory stages by the bulk-async copy engine (cp.async.bulk, SASS UBLKCP,
// completion on an mbarrier); a stage is RB whole lattice rows.  The row above the strip
// is pre-loaded into a carry buffer, afterwards the last row of the previous stage serves
// as the carry, so every byte of the sample is read from HBM exactly once (+ one halo row
// per strip).  HBM-bound by design: algorithmic bytes per sample = n * sizeof(T).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "engine.hpp"

namespace mclcar {
namespace {

constexpr int kThreads = 256;
constexpr int kStages = 4;
constexpr int kMaxNC = 4;  // non-constant design columns handled by the fast kernel

struct StatsTorusParams {
    const void* M;
    int64_t ld, N;
    int n1, n2;
    int RB;           // rows per stage
    int S;            // strips per sample
    int rows_per_strip;
    int64_t items;
    // design: nc non-constant columns (device pointers, [n]) and their WX
    int nc;
    const double* Xc[kMaxNC];
    const double* WXc[kMaxNC];
    int dst_a[kMaxNC], dst_b[kMaxNC];  // statistic row each accumulator goes to
    // constant columns: a_l = cval*sum(x), b_l = 4*cval*sum(x)
    int nconst;
    int cdst_a[MCLCAR_MAX_COV], cdst_b[MCLCAR_MAX_COV];
    double cval[MCLCAR_MAX_COV];
    double* out;  // S == 1: q [d][ldq];  S > 1: partial [S][d][ldq]
    int64_t ldq;
    int d;
};

template <typename T> struct Vec;
template <> struct Vec<double> { static constexpr int N = 2; using type = double2; };
template <> struct Vec<float> { static constexpr int N = 4; using type = float4; };

template <typename T> __device__ __forceinline__ void unpack(const typename Vec<T>::type& v, double (&x)[Vec<T>::N]);
template <> __device__ __forceinline__ void unpack<double>(const double2& v, double (&x)[2]) { x[0] = v.x; x[1] = v.y; }
template <> __device__ __forceinline__ void unpack<float>(const float4& v, double (&x)[4]) {
    x[0] = (double)v.x; x[1] = (double)v.y; x[2] = (double)v.z; x[3] = (double)v.w;
}

template <typename T, int NC>
__global__ void __launch_bounds__(kThreads, 2) stats_torus_kernel(const StatsTorusParams p) {
    constexpr int VEC = Vec<T>::N;
    using VT = typename Vec<T>::type;
    constexpr int NV = 3 + 2 * NC;  // xx, xWx/2, sum x, a[NC], b[NC]

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n2 = p.n2, n1 = p.n1;
    const uint32_t row_bytes = (uint32_t)n2 * sizeof(T);
    const uint32_t stage_bytes = (row_bytes * p.RB + 127u) & ~127u;
    const uint32_t carry_bytes = (row_bytes + 127u) & ~127u;
    T* carry = reinterpret_cast<T*>(smem_raw);
    unsigned char* stage_base = smem_raw + carry_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + (size_t)kStages * stage_bytes);  // [kStages + 1]
    double* scratch = reinterpret_cast<double*>(bars + kStages + 2);

    if (threadIdx.x == 0) {
        for (int b = 0; b <= kStages; ++b) mbar_init(&bars[b], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;  // bit b = parity to wait for on barrier b (uniform over the CTA)

    const int nvr = n2 / VEC;  // vectors per row
    // A thread walks DOWN a column vector through a block of consecutive rows of the stage and keeps the row
    // above in registers: every element is loaded from shared memory and converted to fp64 once (plus the
    // first element of the next vector, and the row above its block: ~1.4 conversions per fp32 element
    // instead of 2.25 when every vector re-read its upper neighbour -- the conversion unit, not HBM, was
    // the limit of the fp32 path).  Rows narrower than the CTA are shared by `groups` row blocks.
    const bool wide = nvr > kThreads;
    const int groups = wide ? 1 : min(kThreads / nvr, p.RB);
    const int cv_t = wide ? (int)threadIdx.x : (int)threadIdx.x % nvr;
    const int rg = wide ? 0 : (int)threadIdx.x / nvr;
    const int cstride = wide ? kThreads : nvr;

    for (int64_t item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int64_t i = item / p.S;
        const int strip = (int)(item - i * p.S);
        const int r0 = strip * p.rows_per_strip;
        const int r1 = min(n1, r0 + p.rows_per_strip);
        const int nrows = r1 - r0;
        const int nst = (nrows + p.RB - 1) / p.RB;
        const T* src = reinterpret_cast<const T*>(p.M) + i * p.ld;

        if (threadIdx.x == 0) {
            const int rc = (r0 + n1 - 1) % n1;
            mbar_expect_tx(&bars[kStages], row_bytes);
            bulk_g2s(carry, src + (size_t)rc * n2, row_bytes, &bars[kStages]);
            for (int s = 0; s < kStages - 1 && s < nst; ++s) {
                const int rs = r0 + s * p.RB;
                const uint32_t bytes = row_bytes * (uint32_t)min(p.RB, r1 - rs);
                mbar_expect_tx(&bars[s], bytes);
                bulk_g2s(stage_base + (size_t)s * stage_bytes, src + (size_t)rs * n2, bytes, &bars[s]);
            }
        }
        double acc[NV];
#pragma unroll
        for (int a = 0; a < NV; ++a) acc[a] = 0.0;

        mbar_wait(&bars[kStages], (phase >> kStages) & 1u);
        phase ^= 1u << kStages;

        for (int it = 0; it < nst; ++it) {
            const int b = it % kStages;
            mbar_wait(&bars[b], (phase >> b) & 1u);
            phase ^= 1u << b;
            const T* buf = reinterpret_cast<const T*>(stage_base + (size_t)b * stage_bytes);
            const T* prev = it == 0 ? carry
                                    : reinterpret_cast<const T*>(stage_base + (size_t)((it - 1) % kStages) * stage_bytes) +
                                          (size_t)(p.RB - 1) * n2;
            const int rs = r0 + it * p.RB;
            const int rows_here = min(p.RB, r1 - rs);
            const int rpg = (rows_here + groups - 1) / groups;
            const int rb0 = rg * rpg, rb1 = min(rows_here, rb0 + rpg);
            if (rg < groups && rb0 < rb1) {
                for (int cv = cv_t; cv < nvr; cv += cstride) {
                    const int c = cv * VEC;
                    const int cn = (cv + 1 == nvr) ? 0 : c + VEC;
                    double up[VEC];
                    unpack<T>(*reinterpret_cast<const VT*>((rb0 > 0 ? buf + (size_t)(rb0 - 1) * n2 : prev) + c), up);
                    for (int row = rb0; row < rb1; ++row) {
                        const T* rowp = buf + (size_t)row * n2;
                        double x[VEC];
                        unpack<T>(*reinterpret_cast<const VT*>(rowp + c), x);
                        const double xr = (double)rowp[cn];
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            const double right = (e + 1 < VEC) ? x[(e + 1) % VEC] : xr;
                            acc[0] = fma(x[e], x[e], acc[0]);
                            acc[1] = fma(x[e], right + up[e], acc[1]);
                            acc[2] += x[e];
                        }
                        if constexpr (NC > 0) {
                            const size_t site = (size_t)(rs + row) * n2 + c;
#pragma unroll
                            for (int l = 0; l < NC; ++l) {
                                if (l < p.nc) {
#pragma unroll
                                    for (int e = 0; e < VEC; ++e) {
                                        acc[3 + l] = fma(__ldg(p.Xc[l] + site + e), x[e], acc[3 + l]);
                                        acc[3 + NC + l] = fma(__ldg(p.WXc[l] + site + e), x[e], acc[3 + NC + l]);
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int e = 0; e < VEC; ++e) up[e] = x[e];
                    }
                }
            }
            __syncthreads();  // stage it consumed; stage it-1 (the carry) is no longer needed
            if (threadIdx.x == 0) {
                const int s = it + kStages - 1;
                if (s < nst) {
                    const int rs2 = r0 + s * p.RB;
                    const uint32_t bytes = row_bytes * (uint32_t)min(p.RB, r1 - rs2);
                    const int bb = s % kStages;  // the buffer consumed at iteration it-1 (unused so far when it == 0)
                    mbar_expect_tx(&bars[bb], bytes);
                    bulk_g2s(stage_base + (size_t)bb * stage_bytes, src + (size_t)rs2 * n2, bytes, &bars[bb]);
                }
            }
        }
        block_sum<NV>(acc, scratch);
        if (threadIdx.x == 0) {
            double* o = p.out + (size_t)strip * p.d * p.ldq + i;
            o[0] = acc[0];
            o[(size_t)1 * p.ldq] = 2.0 * acc[1];
#pragma unroll
            for (int l = 0; l < NC; ++l) {
                if (l < p.nc) {
                    o[(size_t)p.dst_a[l] * p.ldq] = acc[3 + l];
                    o[(size_t)p.dst_b[l] * p.ldq] = acc[3 + NC + l];
                }
            }
            for (int l = 0; l < p.nconst; ++l) {
                o[(size_t)p.cdst_a[l] * p.ldq] = p.cval[l] * acc[2];
                o[(size_t)p.cdst_b[l] * p.ldq] = 4.0 * p.cval[l] * acc[2];
            }
        }
        __syncthreads();
    }
}

// sum the S strip partials of every statistic in strip order
__global__ void strip_sum_kernel(const double* partial, double* q, int S, int d, int64_t ldq, int64_t N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int j = 0; j < d; ++j) {
        double a = 0.0;
        for (int s = 0; s < S; ++s) a += partial[((size_t)s * d + j) * ldq + i];
        q[(size_t)j * ldq + i] = a;
    }
}

// Linear forms X'y and (WX)'y for NON-CONSTANT design columns, a group of samples at a time.  In the streaming kernel above
// every sample re-reads the column and its W-image (16 B per site per column against the sample's 4 or 8): the pass
// becomes L2-bound at a quarter of the HBM rate (config 2, p = 2: 1.6 TB/s).  Here a CTA takes GS samples and a run of
// sites: a thread reads its sites' column values ONCE and applies them to the GS samples (GS independent 16-byte loads in
// flight), so the columns cost 16 / GS bytes per site and sample.  This is a second read of the sample matrix -- the
// single-pass form (a group of samples per ring stage of the streaming kernel) is the next step.
struct LinFormsParams {
    const void* M;
    int64_t ld, N;
    int64_t n;          // sites per sample
    int64_t chunk;      // sites per CTA (multiple of the vector width)
    int nc;
    const double* Xc[kMaxNC];
    const double* WXc[kMaxNC];
    double* out;        // chunks == 1: q rows directly; else partial [chunks][2 nc][ldq]
    int64_t ldq;
    int dst_a[kMaxNC], dst_b[kMaxNC];
    int chunks;
};

template <typename T, int NC, int GS>
__global__ void __launch_bounds__(256) stats_linear_forms_kernel(const LinFormsParams p) {
    constexpr int VEC = Vec<T>::N;
    using VT = typename Vec<T>::type;
    __shared__ double scratch[8 * GS * 2 * NC];
    const int64_t i0 = (int64_t)blockIdx.y * GS;
    const int gs = (int)min((int64_t)GS, p.N - i0);
    const int64_t s0 = (int64_t)blockIdx.x * p.chunk, s1 = min(p.n, s0 + p.chunk);
    double acc[GS][2 * NC];
#pragma unroll
    for (int g = 0; g < GS; ++g)
#pragma unroll
        for (int l = 0; l < 2 * NC; ++l) acc[g][l] = 0.0;
    const T* base = reinterpret_cast<const T*>(p.M) + i0 * p.ld;
    for (int64_t s = s0 + (int64_t)threadIdx.x * VEC; s < s1; s += (int64_t)blockDim.x * VEC) {
        double xc[NC][VEC], wc[NC][VEC];
#pragma unroll
        for (int l = 0; l < NC; ++l)
#pragma unroll
            for (int e = 0; e < VEC; e += 2) {
                const double2 a = __ldg(reinterpret_cast<const double2*>(p.Xc[l] + s + e));
                const double2 b = __ldg(reinterpret_cast<const double2*>(p.WXc[l] + s + e));
                xc[l][e] = a.x; xc[l][e + 1] = a.y; wc[l][e] = b.x; wc[l][e + 1] = b.y;
            }
        VT v[GS];
#pragma unroll
        for (int g = 0; g < GS; ++g)
            if (g < gs) v[g] = *reinterpret_cast<const VT*>(base + (int64_t)g * p.ld + s);
#pragma unroll
        for (int g = 0; g < GS; ++g) {
            if (g < gs) {
                double x[VEC];
                unpack<T>(v[g], x);
#pragma unroll
                for (int l = 0; l < NC; ++l)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        acc[g][l] = fma(xc[l][e], x[e], acc[g][l]);
                        acc[g][NC + l] = fma(wc[l][e], x[e], acc[g][NC + l]);
                    }
            }
        }
    }
    // fixed-order reduction: lanes by shuffles, the eight warps through shared memory
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int g = 0; g < GS; ++g)
#pragma unroll
        for (int l = 0; l < 2 * NC; ++l) {
            const double t = warp_sum(acc[g][l]);
            if (lane == 0) scratch[(wid * GS + g) * 2 * NC + l] = t;
        }
    __syncthreads();
    if (threadIdx.x < GS * 2 * NC) {
        const int g = threadIdx.x / (2 * NC), l = threadIdx.x % (2 * NC);
        if (g < gs) {
            double a = 0.0;
            for (int w = 0; w < 8; ++w) a += scratch[(w * GS + g) * 2 * NC + l];
            const int row = l < NC ? p.dst_a[l] : p.dst_b[l - NC];
            if (p.chunks == 1) p.out[(size_t)row * p.ldq + i0 + g] = a;
            else p.out[((size_t)blockIdx.x * 2 * NC + l) * p.ldq + i0 + g] = a;
        }
    }
}

// partial [chunks][2 nc][ldq] -> rows dst_a / dst_b of q, chunks added in order
__global__ void linear_forms_sum_kernel(const LinFormsParams p, double* q) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    for (int l = 0; l < 2 * p.nc; ++l) {
        double a = 0.0;
        for (int c = 0; c < p.chunks; ++c) a += p.out[((size_t)c * 2 * p.nc + l) * p.ldq + i];
        const int row = l < p.nc ? p.dst_a[l] : p.dst_b[l - p.nc];
        q[(size_t)row * p.ldq + i] = a;
    }
}

template <typename T>
int launch_linear_forms(mclcar_ctx* ctx, LinFormsParams& p, double* d_q) {
    constexpr int VEC = Vec<T>::N;
    const int GS = p.nc <= 2 ? 8 : 4;
    const int64_t groups = (p.N + GS - 1) / GS;
    // sites per CTA: whole lattice when the sample groups alone fill the machine, else split (partials added in order)
    int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)4 * ctx->n_sm + groups - 1) / groups, p.n / (256 * VEC * 4) + 1));
    if (chunks > 65535) chunks = 65535;
    int64_t chunk = (p.n + chunks - 1) / chunks;
    chunk = (chunk + VEC - 1) / VEC * VEC;
    chunks = (int)((p.n + chunk - 1) / chunk);
    p.chunk = chunk; p.chunks = chunks;
    double* partial = nullptr;
    if (chunks > 1) {
        MCL_TRY(pool_alloc(ctx, (void**)&partial, (size_t)chunks * 2 * p.nc * p.ldq * sizeof(double)));
        p.out = partial;
    } else {
        p.out = d_q;
    }
    if (groups > 65535) { pool_free(ctx, partial); return fail(ctx, MCLCAR_ELIMIT, "too many samples for the linear-forms kernel"); }
    dim3 grid((unsigned)chunks, (unsigned)groups);
    cudaStream_t st = ctx->stream;
    switch (p.nc) {
        case 1: stats_linear_forms_kernel<T, 1, 8><<<grid, 256, 0, st>>>(p); break;
        case 2: stats_linear_forms_kernel<T, 2, 8><<<grid, 256, 0, st>>>(p); break;
        case 3: stats_linear_forms_kernel<T, 3, 4><<<grid, 256, 0, st>>>(p); break;
        default: stats_linear_forms_kernel<T, 4, 4><<<grid, 256, 0, st>>>(p); break;
    }
    if (chunks > 1) linear_forms_sum_kernel<<<(unsigned)((p.N + 255) / 256), 256, 0, st>>>(p, d_q);
    cudaError_t e = cudaGetLastError();
    pool_free(ctx, partial);    // stream-ordered reuse: later work of this context is queued behind the kernels
    if (e != cudaSuccess) return fail(ctx, MCLCAR_ECUDA, "linear-forms kernel: %s", cudaGetErrorString(e));
    return MCLCAR_OK;
}

template <typename T>
int launch_typed(mclcar_ctx* ctx, const StatsTorusParams& p, int grid, size_t smem) {
    cudaStream_t st = ctx->stream;
#define LAUNCH(NC)                                                                                              \
    do {                                                                                                        \
        MCL_CUDA(ctx, cudaFuncSetAttribute(stats_torus_kernel<T, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem));                                                          \
        stats_torus_kernel<T, NC><<<grid, kThreads, smem, st>>>(p);                                              \
    } while (0)
    if (p.nc == 0) LAUNCH(0);
    else if (p.nc == 1) LAUNCH(1);
    else if (p.nc == 2) LAUNCH(2);
    else LAUNCH(4);
#undef LAUNCH
    MCL_CUDA(ctx, cudaGetLastError());
    return MCLCAR_OK;
}

}  // namespace

// returns MCLCAR_OK and sets *handled = false when the shape does not fit the fast path
int launch_stats_torus_fast(mclcar_ctx* ctx, const void* M, int dtype, int64_t N, int64_t ld, double* d_q,
                            int64_t ldq, bool* handled) {
    *handled = false;
    MCL_TRY(ensure_algebra(ctx));   // d_WX
    const int n1 = ctx->n1, n2 = ctx->n2, p = ctx->p;
    const size_t es = dtype == MCLCAR_F32 ? 4 : 8;
    const int vec = dtype == MCLCAR_F32 ? 4 : 2;
    const size_t row_bytes = (size_t)n2 * es;
    if (n2 % vec != 0 || row_bytes % 16 != 0 || row_bytes > 24 * 1024) return MCLCAR_OK;
    if (((uintptr_t)M % 16) != 0 || ((size_t)ld * es) % 16 != 0) return MCLCAR_OK;
    StatsTorusParams prm{};
    prm.M = M; prm.ld = ld; prm.N = N; prm.n1 = n1; prm.n2 = n2;
    // design columns: constant columns need no memory traffic on the torus (WX = 4 X)
    prm.nc = 0; prm.nconst = 0;
    for (int l = 0; l < p; ++l) {
        const double* col = ctx->X.data() + (size_t)l * ctx->n;
        bool is_const = true;
        for (int i = 1; i < ctx->n && is_const; ++i) is_const = col[i] == col[0];
        if (is_const) {
            prm.cdst_a[prm.nconst] = 2 + l; prm.cdst_b[prm.nconst] = 2 + p + l; prm.cval[prm.nconst] = col[0];
            ++prm.nconst;
        } else {
            if (prm.nc == kMaxNC) return MCLCAR_OK;
            prm.Xc[prm.nc] = ctx->d_X + (size_t)l * ctx->n; prm.WXc[prm.nc] = ctx->d_WX + (size_t)l * ctx->n;
            prm.dst_a[prm.nc] = 2 + l; prm.dst_b[prm.nc] = 2 + p + l;
            ++prm.nc;
        }
    }
    // stage geometry: <= 24 KB per stage, 4 stages + carry, two CTAs per SM
    const size_t stage_target = 24 * 1024;
    int RB = (int)(stage_target / row_bytes);
    if (RB < 1) RB = 1;
    if (RB > n1) RB = n1;
    prm.RB = RB;
    // strips: enough work items to fill the machine twice
    const int64_t want = (int64_t)4 * ctx->n_sm;
    int S = 1;
    if (N < want) {
        S = (int)((want + N - 1) / N);
        const int max_s = (n1 + RB - 1) / RB;
        if (S > max_s) S = max_s;
        if (S < 1) S = 1;
    }
    prm.rows_per_strip = (n1 + S - 1) / S;
    S = (n1 + prm.rows_per_strip - 1) / prm.rows_per_strip;
    prm.S = S;
    prm.items = N * S;
    prm.d = 2 + 2 * p; prm.ldq = ldq;
    double* partial = nullptr;
    if (S > 1) {
        MCL_TRY(ensure(ctx, ctx->d_partial, (size_t)S * prm.d * ldq * sizeof(double)));
        partial = (double*)ctx->d_partial.p;
    }
    prm.out = S > 1 ? partial : d_q;

    const size_t stage_bytes = (row_bytes * RB + 127) & ~(size_t)127;
    const size_t carry_bytes = (row_bytes + 127) & ~(size_t)127;
    const size_t smem = carry_bytes + kStages * stage_bytes + (kStages + 2) * sizeof(uint64_t) +
                        (kThreads / 32) * (3 + 2 * kMaxNC) * sizeof(double) + 128;
    int grid = (int)std::min<int64_t>(prm.items, (int64_t)2 * ctx->n_sm);
    // non-constant design columns: the quadratic forms and the constant columns in the streaming pass, the linear forms
    // of the others in the grouped kernel (MCLCAR_STATS_LINEAR=0 keeps them in the streaming pass)
    const char* le = getenv("MCLCAR_STATS_LINEAR");
    const bool split = prm.nc > 0 && !(le && atoi(le) == 0) && (ld % vec) == 0 && (ctx->n % vec) == 0;
    LinFormsParams lf{};
    if (split) {
        lf.M = M; lf.ld = ld; lf.N = N; lf.n = ctx->n; lf.nc = prm.nc; lf.ldq = ldq;
        for (int l = 0; l < prm.nc; ++l) { lf.Xc[l] = prm.Xc[l]; lf.WXc[l] = prm.WXc[l]; lf.dst_a[l] = prm.dst_a[l]; lf.dst_b[l] = prm.dst_b[l]; }
        prm.nc = 0;
    }
    ProfScope prof(ctx, PROF_STATS, (S > 1 ? 2 : 1) + (split ? 1 : 0));
    if (dtype == MCLCAR_F32) MCL_TRY(launch_typed<float>(ctx, prm, grid, smem));
    else MCL_TRY(launch_typed<double>(ctx, prm, grid, smem));
    if (S > 1) {
        strip_sum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(partial, d_q, S, prm.d, ldq, N);
        MCL_CUDA(ctx, cudaGetLastError());
    }
    if (split) {   // after the strip sums: they write every row of q
        if (dtype == MCLCAR_F32) MCL_TRY(launch_linear_forms<float>(ctx, lf, d_q));
        else MCL_TRY(launch_linear_forms<double>(ctx, lf, d_q));
    }
    *handled = true;
    return MCLCAR_OK;
}

}  // namespace mclcar

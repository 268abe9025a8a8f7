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
    ProfScope prof(ctx, PROF_STATS, S > 1 ? 2 : 1);
    if (dtype == MCLCAR_F32) MCL_TRY(launch_typed<float>(ctx, prm, grid, smem));
    else MCL_TRY(launch_typed<double>(ctx, prm, grid, smem));
    if (S > 1) {
        strip_sum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(partial, d_q, S, prm.d, ldq, N);
        MCL_CUDA(ctx, cudaGetLastError());
    }
    *handled = true;
    return MCLCAR_OK;
}

}  // namespace mclcar

// K1 (torus) -- batched red-black Gibbs sampler for the zero-mean CAR field
//     z ~ N(0, sigma (I - rho W)^-1),   z_s | rest ~ N(rho * sum_{j~s} z_j, sigma)
// on the rook torus of CAR.simTorus (R/CAR.simLM.R:2-22).  It replaces the reference's
// exact draw (a fresh sparse Cholesky + backsolve per sample, R/CAR.simLM.R:28-31); its RNG
// stream (Philox4x32-10) differs from R's, so it is validated statistically.
//
// Layout: every chain keeps its lattice as two half-lattices ("red": (i + j) even, "black":
// (i + j) odd), each n1 x n2/2 row-major, so that a half sweep reads only the other colour
// and writes only its own with unit-stride 16-byte accesses: 2 * sizeof(float) = 8 bytes of
// HBM traffic per site update.  Site (i, j) of colour col sits at half-column c with
// j = 2c + ((i & 1) ^ col); its four neighbours are other[i-1][c], other[i+1][c], other[i][c]
// and other[i][c + s], s = +1 if ((i & 1) ^ col) else -1 (cyclic).
// One thread updates 4 consecutive half-columns of 2 consecutive rows, each 4-vector from ONE
// Philox4x32-10 call keyed by (seed; half-lattice vector index, sweep*2+colour, global chain
// id) -- results do not depend on the launch geometry nor on how chains are sharded over GPUs.
//
// Over-relaxation (Adler 1981; Fox & Parker 2017 for the sampler = solver equivalence): the
// update x' = m + (1-omega)(x - m) + sqrt(omega(2-omega) sigma) z leaves N(m, sigma) invariant
// for every omega in (0, 2); with the red-black ordering and omega = 2/(1 + sqrt(1 - mu^2)),
// mu = |rho| lambda_max(W), the chain decorrelates at rate omega - 1 per sweep instead of
// mu^2 (0.25 instead of 0.64 at rho = 0.2).  omega = 1 is the plain Gibbs sampler.
#include "common.cuh"
#include "engine.hpp"

namespace mclcar {
namespace {

constexpr uint32_t kTagTorus = 0x7052u;
constexpr int kRows = 2;  // lattice rows per thread

struct TorusSweepParams {
    float* red;    // [C][n1][h]
    float* black;  // [C][n1][h]
    int n1, h;     // h = n2 / 2
    int64_t C;
    // over-relaxed Gibbs update  x' = keep*x + gain*(sum of neighbours) + noise*z  with
    // keep = 1 - omega, gain = omega*rho, noise = sqrt(omega (2 - omega) sigma); omega = 1 is plain Gibbs
    float keep, gain, noise;
    uint32_t sweep;
    int64_t chain0, chain_stride;  // global chain id = chain0 + local * chain_stride
    PhiloxKeys keys;
};

// VEC = 4: h % 4 == 0, 16-byte path.  VEC = 1: any h.  SOR: read the current value too.
template <int VEC, bool SOR>
__global__ void __launch_bounds__(256) gibbs_torus_half_sweep(const TorusSweepParams p, const int col) {
    const int h = p.h, n1 = p.n1;
    const int hv = h / VEC;
    const int cv = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = (blockIdx.y * blockDim.y + threadIdx.y) * kRows;
    const int64_t ch = blockIdx.z;
    if (cv >= hv || i0 >= n1) return;
    const size_t plane = (size_t)n1 * h;
    float* __restrict__ own = (col == 0 ? p.red : p.black) + ch * plane;
    const float* __restrict__ oth = (col == 0 ? p.black : p.red) + ch * plane;
    const uint64_t gch = (uint64_t)(p.chain0 + ch * p.chain_stride);
    const uint32_t c2 = (uint32_t)gch, c3 = (uint32_t)(gch >> 32) ^ (kTagTorus << 16), c1 = p.sweep * 2u + (uint32_t)col;
    const int c0 = cv * VEC;

    if constexpr (VEC == 4) {
        const int cl = c0 == 0 ? h - 1 : c0 - 1, cr = c0 + 4 == h ? 0 : c0 + 4;
        int ip = i0 == 0 ? n1 - 1 : i0 - 1;
        float4 up = *reinterpret_cast<const float4*>(oth + (unsigned)(ip * h + c0));
        float4 o = *reinterpret_cast<const float4*>(oth + (unsigned)(i0 * h + c0));
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
            const int i = i0 + k;
            if (i >= n1) break;
            const int in = i + 1 == n1 ? 0 : i + 1;
            const float4 d = *reinterpret_cast<const float4*>(oth + (unsigned)(in * h + c0));
            const bool fwd = ((i & 1) ^ col) != 0;  // same-row neighbour sits at c + 1, else at c - 1
            const float e = oth[(unsigned)(i * h + (fwd ? cr : cl))];
            const Philox4 r = philox4x32_10((uint32_t)(i * hv + cv), c1, c2, c3, p.keys);
            float z0, z1, z2, z3;
            box_muller(r.x, r.y, z0, z1);
            box_muller(r.z, r.w, z2, z3);
            float4 s;
            if (fwd) {
                s.x = o.x + o.y; s.y = o.y + o.z; s.z = o.z + o.w; s.w = o.w + e;
            } else {
                s.x = e + o.x; s.y = o.x + o.y; s.z = o.y + o.z; s.w = o.z + o.w;
            }
            float4 v;
            v.x = fmaf(p.gain, s.x + (up.x + d.x), p.noise * z0);
            v.y = fmaf(p.gain, s.y + (up.y + d.y), p.noise * z1);
            v.z = fmaf(p.gain, s.z + (up.z + d.z), p.noise * z2);
            v.w = fmaf(p.gain, s.w + (up.w + d.w), p.noise * z3);
            float4* dst = reinterpret_cast<float4*>(own + (unsigned)(i * h + c0));
            if constexpr (SOR) {
                const float4 cur = *dst;
                v.x = fmaf(p.keep, cur.x, v.x);
                v.y = fmaf(p.keep, cur.y, v.y);
                v.z = fmaf(p.keep, cur.z, v.z);
                v.w = fmaf(p.keep, cur.w, v.w);
            }
            *dst = v;
            up = o;
            o = d;
        }
    } else {
        for (int k = 0; k < kRows; ++k) {
            const int i = i0 + k;
            if (i >= n1) break;
            const int iu = i == 0 ? n1 - 1 : i - 1, id = i + 1 == n1 ? 0 : i + 1;
            const bool fwd = ((i & 1) ^ col) != 0;
            const int c = c0;
            const int cn = fwd ? (c + 1 == h ? 0 : c + 1) : (c == 0 ? h - 1 : c - 1);
            const Philox4 r = philox4x32_10((uint32_t)(i * hv + cv), c1, c2, c3, p.keys);
            float z0, z1;
            box_muller(r.x, r.y, z0, z1);
            const float s = oth[(size_t)i * h + c] + oth[(size_t)i * h + cn] + oth[(size_t)iu * h + c] + oth[(size_t)id * h + c];
            float v = fmaf(p.gain, s, p.noise * z0);
            if constexpr (SOR) v = fmaf(p.keep, own[(size_t)i * h + c], v);
            own[(size_t)i * h + c] = v;
        }
    }
}

// write the current state of chains [0, nch) as samples: out[(slot0 + ch)*ld + site] = mu[site] + z
template <typename TO>
__global__ void __launch_bounds__(256) torus_emit_kernel(const float* red, const float* black, int n1, int h,
                                                         const double* mu, TO* out, int64_t ld, int64_t slot0,
                                                         int64_t nch) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    const int64_t ch = blockIdx.z;
    if (c >= h || ch >= nch) return;
    const size_t plane = (size_t)n1 * h;
    const float r = red[ch * plane + (size_t)i * h + c];
    const float b = black[ch * plane + (size_t)i * h + c];
    const int par = i & 1;  // red sits at column 2c + par, black at 2c + 1 - par
    const size_t site = (size_t)i * 2 * h + 2 * c;
    const double m0 = mu ? mu[site] : 0.0, m1 = mu ? mu[site + 1] : 0.0;
    const double v0 = (par ? (double)b : (double)r) + m0;
    const double v1 = (par ? (double)r : (double)b) + m1;
    TO* o = out + (size_t)(slot0 + ch) * ld + site;
    o[0] = (TO)v0;
    o[1] = (TO)v1;
}

}  // namespace

int torus_chains_alloc(mclcar_ctx* ctx, int64_t C, TorusChains& t) {
    const size_t plane = (size_t)ctx->n1 * (ctx->n2 / 2);
    MCL_TRY(pool_alloc(ctx, (void**)&t.red, (size_t)C * plane * sizeof(float)));
    MCL_TRY(pool_alloc(ctx, (void**)&t.black, (size_t)C * plane * sizeof(float)));
    MCL_CUDA(ctx, cudaMemsetAsync(t.red, 0, (size_t)C * plane * sizeof(float), ctx->stream));
    MCL_CUDA(ctx, cudaMemsetAsync(t.black, 0, (size_t)C * plane * sizeof(float), ctx->stream));
    t.C = C;
    return MCLCAR_OK;
}

void torus_chains_free(mclcar_ctx* ctx, TorusChains& t) {
    pool_free(ctx, t.red);
    pool_free(ctx, t.black);
    t.red = t.black = nullptr;
}

// one full sweep (both colours) of all chains; omega in [1, 2): over-relaxation factor
int torus_sweep(mclcar_ctx* ctx, TorusChains& t, double rho, double sigma, double omega, uint32_t sweep) {
    TorusSweepParams p{};
    p.red = t.red; p.black = t.black; p.n1 = ctx->n1; p.h = ctx->n2 / 2; p.C = t.C;
    p.keep = (float)(1.0 - omega);
    p.gain = (float)(omega * rho);
    p.noise = (float)std::sqrt(omega * (2.0 - omega) * sigma);
    p.keys = philox_keys(ctx->seed);
    p.sweep = sweep;
    p.chain0 = ctx->rank; p.chain_stride = ctx->world;
    const bool vec4 = (p.h % 4) == 0;
    const bool sor = omega != 1.0;
    const int hv = vec4 ? p.h / 4 : p.h;
    dim3 block(32, 8);
    while ((int)block.x / 2 >= hv && block.x > 1) { block.x /= 2; block.y *= 2; }
    const int rows_t = (p.n1 + kRows - 1) / kRows;
    dim3 grid((hv + block.x - 1) / block.x, (rows_t + block.y - 1) / block.y, (unsigned)t.C);
    if (t.C > 65535) return fail(ctx, MCLCAR_ELIMIT, "more than 65535 chains in one torus sampler launch");
    if ((int64_t)p.n1 * p.h >= ((int64_t)1 << 31)) return fail(ctx, MCLCAR_ELIMIT, "lattice too large for 32-bit indexing");
    ProfScope prof(ctx, PROF_SWEEP, 2);
    for (int col = 0; col < 2; ++col) {
        if (vec4 && sor) gibbs_torus_half_sweep<4, true><<<grid, block, 0, ctx->stream>>>(p, col);
        else if (vec4) gibbs_torus_half_sweep<4, false><<<grid, block, 0, ctx->stream>>>(p, col);
        else if (sor) gibbs_torus_half_sweep<1, true><<<grid, block, 0, ctx->stream>>>(p, col);
        else gibbs_torus_half_sweep<1, false><<<grid, block, 0, ctx->stream>>>(p, col);
    }
    MCL_CUDA(ctx, cudaGetLastError());
    return MCLCAR_OK;
}

int torus_emit(mclcar_ctx* ctx, TorusChains& t, const double* d_mu, void* out, int dtype, int64_t ld, int64_t slot0,
               int64_t nch) {
    const int h = ctx->n2 / 2;
    dim3 block(128);
    dim3 grid((h + 127) / 128, ctx->n1, (unsigned)nch);
    ProfScope prof(ctx, PROF_EMIT, 1);
    if (dtype == MCLCAR_F32)
        torus_emit_kernel<float><<<grid, block, 0, ctx->stream>>>(t.red, t.black, ctx->n1, h, d_mu, (float*)out, ld, slot0, nch);
    else
        torus_emit_kernel<double><<<grid, block, 0, ctx->stream>>>(t.red, t.black, ctx->n1, h, d_mu, (double*)out, ld, slot0, nch);
    MCL_CUDA(ctx, cudaGetLastError());
    return MCLCAR_OK;
}

}  // namespace mclcar

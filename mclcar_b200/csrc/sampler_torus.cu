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
// One thread updates 4 consecutive half-columns of 8 consecutive rows, each 4-vector from ONE
// Philox4x32-10 call keyed by (seed; half-lattice vector index, sweep*2+colour, global chain
// id) -- results do not depend on the launch geometry nor on how chains are sharded over GPUs.
//
// Over-relaxation (Adler 1981; Fox & Parker 2017 for the sampler = solver equivalence): the
// update x' = m + (1-omega)(x - m) + sqrt(omega(2-omega) sigma) z leaves N(m, sigma) invariant
// for every omega in (0, 2); with the red-black ordering and omega = 2/(1 + sqrt(1 - mu^2)),
// mu = |rho| lambda_max(W), the chain decorrelates at rate omega - 1 per sweep instead of
// mu^2 (0.25 instead of 0.64 at rho = 0.2).  omega = 1 is the plain Gibbs sampler.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "engine.hpp"

namespace mclcar {
namespace {

constexpr uint32_t kTagTorus = 0x7052u;

struct TorusSweepParams {
    float* red;    // [C][n1][h]
    float* black;  // [C][n1][h]
    float* red_out;    // fused sweep only: the other buffer of the ping-pong pair
    float* black_out;
    int n1, h;     // h = n2 / 2
    int64_t C;
    // over-relaxed Gibbs update  x' = keep*x + gain*(sum of neighbours) + noise*z  with
    // keep = 1 - omega, gain = omega*rho, noise = sqrt(omega (2 - omega) sigma); omega = 1 is plain Gibbs
    float keep, gain, noise;
    float rad2;    // -2 ln2 * noise^2: radius of the scaled Box-Muller pair = sqrt(rad2 * lg2(u))
    uint32_t sweep;
    int64_t chain0, chain_stride;  // global chain id = chain0 + local * chain_stride
    PhiloxKeys keys;
    // emission (black half sweep of the last sweep of a thinning block): sample = mu + state,
    // written in the reference's site order to out[(slot0 + chain) * ld + site]
    void* out;
    int64_t ld, slot0, nch;
    const double* mu;   // [n] or null
    double mu_const;    // used when mu is null
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// The update of four consecutive half-columns of one row (ONE Philox4x32-10 block), shared by the
// two-launch and the fused kernels so that both produce the same bits:
//     x' = keep * x + gain * (sum of the four neighbours) + noise * z,   z ~ N(0, 1).
// o/up/d: the other colour at the same half-columns in this row / the row above / below; e: the
// same-row neighbour beyond the vector (c + 4 when FWD, c - 1 otherwise).  The noise scale is folded
// into the Box-Muller radius (sqrt(rad2 * lg2 u), rad2 = -2 ln2 noise^2) and the product
// radius * cos/sin into the final FMA: 6 FP32 instructions per site besides the RNG.
template <bool SOR, bool FWD>
__device__ __forceinline__ float4 gibbs_update4(const float4 o, const float4 up, const float4 d, const float e,
                                                const float4 cur, const Philox4 r, const float keep, const float gain,
                                                const float rad2) {
    const float ra = sqrt_approx(rad2 * lg2_approx(u01(r.x)));
    const float rb = sqrt_approx(rad2 * lg2_approx(u01(r.z)));
    float sa, ca, sb, cb;
    sincos_approx(6.2831853071795865f * u01(r.y), sa, ca);
    sincos_approx(6.2831853071795865f * u01(r.w), sb, cb);
    float4 s;
    if (FWD) { s.x = o.x + o.y; s.y = o.y + o.z; s.z = o.z + o.w; s.w = o.w + e; }
    else { s.x = e + o.x; s.y = o.x + o.y; s.z = o.y + o.z; s.w = o.z + o.w; }
    s.x = __fadd_rn(s.x, __fadd_rn(up.x, d.x));
    s.y = __fadd_rn(s.y, __fadd_rn(up.y, d.y));
    s.z = __fadd_rn(s.z, __fadd_rn(up.z, d.z));
    s.w = __fadd_rn(s.w, __fadd_rn(up.w, d.w));
    float4 v;
    if (SOR) {
        v.x = __fmaf_rn(ra, ca, __fmul_rn(keep, cur.x));
        v.y = __fmaf_rn(ra, sa, __fmul_rn(keep, cur.y));
        v.z = __fmaf_rn(rb, cb, __fmul_rn(keep, cur.z));
        v.w = __fmaf_rn(rb, sb, __fmul_rn(keep, cur.w));
    } else {
        v.x = __fmul_rn(ra, ca); v.y = __fmul_rn(ra, sa); v.z = __fmul_rn(rb, cb); v.w = __fmul_rn(rb, sb);
    }
    v.x = __fmaf_rn(gain, s.x, v.x);
    v.y = __fmaf_rn(gain, s.y, v.y);
    v.z = __fmaf_rn(gain, s.z, v.z);
    v.w = __fmaf_rn(gain, s.w, v.w);
    return v;
}

// 16-byte path (h % 4 == 0).  SOR: the update reads the current value too.  EMIT: the black
// half sweep also stores the finished sample row segment (both colours are in registers).
template <bool SOR, bool EMIT, typename TO, int kRows>
__global__ void __launch_bounds__(256) gibbs_torus_half_sweep_v4(const TorusSweepParams p, const int col) {
    const int h = p.h, n1 = p.n1;
    const int hv = h >> 2;
    const int cv = blockIdx.x * 32 + (threadIdx.x & 31);
    const int i0 = (blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kRows;
    const int64_t ch = blockIdx.z;
    if (cv >= hv || i0 >= n1) return;
    const size_t plane = (size_t)n1 * h;
    float* __restrict__ own = (col == 0 ? p.red : p.black) + ch * plane;
    const float* __restrict__ oth = (col == 0 ? p.black : p.red) + ch * plane;
    const uint64_t gch = (uint64_t)(p.chain0 + ch * p.chain_stride);
    const uint32_t c2 = (uint32_t)gch, c3 = (uint32_t)(gch >> 32) ^ (kTagTorus << 16), c1 = p.sweep * 2u + (uint32_t)col;
    const int c0 = cv << 2;
    const int cl = c0 == 0 ? h - 1 : c0 - 1, cr = c0 + 4 == h ? 0 : c0 + 4;
    const int iend = min(i0 + kRows, n1);

    int row = i0 * h;                                  // offset of row i (floats)
    float4 up = ld4(oth + (i0 == 0 ? (n1 - 1) * h : row - h) + c0);
    float4 o = ld4(oth + row + c0);
    uint32_t ctr = (uint32_t)(i0 * hv + cv);
    // loads of row i are issued one iteration ahead (software prefetch in registers)
    int rown = i0 + 1 == n1 ? 0 : row + h;
    float4 d = ld4(oth + rown + c0);
    float e = oth[row + ((((i0 & 1) ^ col) != 0) ? cr : cl)];
    float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
    if (SOR) cur = ld4(own + row + c0);
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const int i = i0 + k;
        if (i >= iend) break;
        const bool fwd = ((i & 1) ^ col) != 0;          // same-row neighbour sits at c + 1, else at c - 1
        // prefetch row i + 1
        const int row2 = i + 2 >= n1 ? (i + 2 - n1) * h : rown + h;
        float4 dn = d, curn = cur;
        float en = e;
        if (k + 1 < kRows && i + 1 < iend) {
            dn = ld4(oth + row2 + c0);
            en = oth[rown + (fwd ? cl : cr)];           // row i + 1 has the opposite parity
            if (SOR) curn = ld4(own + rown + c0);
        }
        const Philox4 r = philox4x32_10(ctr, c1, c2, c3, p.keys);
        const float4 v = fwd ? gibbs_update4<SOR, true>(o, up, d, e, cur, r, p.keep, p.gain, p.rad2)
                             : gibbs_update4<SOR, false>(o, up, d, e, cur, r, p.keep, p.gain, p.rad2);
        *reinterpret_cast<float4*>(own + row + c0) = v;
        if (EMIT) {
            if (ch < p.nch) {
                // black half sweep (col == 1): o = finished red values, v = new black values of the same
                // half-columns; red sits at column 2c + par, black at 2c + 1 - par (par = i & 1)
                const bool par = (i & 1) != 0;
                const size_t site = (size_t)i * 2 * h + 2 * c0;
                float a[8];
                a[0] = par ? v.x : o.x; a[1] = par ? o.x : v.x; a[2] = par ? v.y : o.y; a[3] = par ? o.y : v.y;
                a[4] = par ? v.z : o.z; a[5] = par ? o.z : v.z; a[6] = par ? v.w : o.w; a[7] = par ? o.w : v.w;
                TO* dst = reinterpret_cast<TO*>(p.out) + (size_t)(p.slot0 + ch) * p.ld + site;
                if (p.mu != nullptr) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) dst[t] = (TO)((double)a[t] + __ldg(p.mu + site + t));
                } else if (sizeof(TO) == 4) {
                    const float m = (float)p.mu_const;
                    float4* d4 = reinterpret_cast<float4*>(dst);
                    d4[0] = make_float4(a[0] + m, a[1] + m, a[2] + m, a[3] + m);
                    d4[1] = make_float4(a[4] + m, a[5] + m, a[6] + m, a[7] + m);
                } else {
#pragma unroll
                    for (int t = 0; t < 8; ++t) dst[t] = (TO)((double)a[t] + p.mu_const);
                }
            }
        }
        up = o;
        o = d;
        d = dn;
        e = en;
        cur = curn;
        row = rown;
        rown = row2;
        ctr += (uint32_t)hv;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused full sweep: ONE pass over HBM per sweep instead of two.  A CTA owns TR lattice rows of
// one chain at full width.  Black rows r0-2 .. r0+TR+1 and red rows r0-1 .. r0+TR are staged into
// shared memory by the bulk-async copy engine (cp.async.bulk, one mbarrier); phase 1 updates red
// for the TR rows AND the two halo rows (recomputing what the neighbouring tiles compute -- same
// Philox counters, hence the same numbers), phase 2 updates black for the TR rows from the new
// red values in shared memory.  Traffic: (2TR+6)/(2TR) reads + 1 write of the state per sweep,
// about 9 B per site update instead of 12 (over-relaxed) for the two-launch form; results are
// bit-identical to it.
// ---------------------------------------------------------------------------------------------
// One lattice row of one colour inside the fused sweep: lanes stride the column vectors.  FWD (the
// same-row neighbour sits at c + 1, else c - 1) is uniform over the row, so it is a template
// argument.  PH2 = black phase: the own colour is only read (current value), the new row goes to
// global memory (and to the sample matrix when EMIT).  The loop body is kept small on purpose: a
// four-way unrolled straight-line variant (more instruction-level parallelism, 4x the code) measured
// 5 % slower -- the 24 resident warps already cover the latencies, instruction fetch does not.
template <bool SOR, bool FWD, bool PH2, bool EMIT, typename TO>
__device__ __forceinline__ void fused_row(const TorusSweepParams& p, const float* __restrict__ nm, float* __restrict__ own,
                                          float* __restrict__ gout, TO* __restrict__ emit_row, const double* __restrict__ mu_row,
                                          const bool par, const int tx, const int hv, const int h, const uint32_t ctr0,
                                          const uint32_t c1, const uint32_t c2, const uint32_t c3) {
    for (int cv = tx; cv < hv; cv += 32) {
        const int c0 = cv << 2;
        const float4 o = ld4(nm + c0), up = ld4(nm - h + c0), d = ld4(nm + h + c0);
        const float e = FWD ? nm[c0 + 4 == h ? 0 : c0 + 4] : nm[c0 == 0 ? h - 1 : c0 - 1];
        float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
        if (SOR) cur = ld4(own + c0);
        const Philox4 r = philox4x32_10(ctr0 + (uint32_t)cv, c1, c2, c3, p.keys);
        const float4 v = gibbs_update4<SOR, FWD>(o, up, d, e, cur, r, p.keep, p.gain, p.rad2);
        if (!PH2) *reinterpret_cast<float4*>(own + c0) = v;
        if (gout) *reinterpret_cast<float4*>(gout + c0) = v;
        if (EMIT && PH2) {
            if (emit_row) {
                // o = finished values of the other colour at the same half-columns; the emitting (black) colour sits at
                // column 2c + 1 - par, the other at 2c + par
                float a[8];
                a[0] = par ? v.x : o.x; a[1] = par ? o.x : v.x; a[2] = par ? v.y : o.y; a[3] = par ? o.y : v.y;
                a[4] = par ? v.z : o.z; a[5] = par ? o.z : v.z; a[6] = par ? v.w : o.w; a[7] = par ? o.w : v.w;
                TO* dst = emit_row + 2 * c0;
                if (mu_row != nullptr) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) dst[t] = (TO)((double)a[t] + __ldg(mu_row + 2 * c0 + t));
                } else if (sizeof(TO) == 4) {
                    const float m = (float)p.mu_const;
                    float4* d4 = reinterpret_cast<float4*>(dst);
                    d4[0] = make_float4(a[0] + m, a[1] + m, a[2] + m, a[3] + m);
                    d4[1] = make_float4(a[4] + m, a[5] + m, a[6] + m, a[7] + m);
                } else {
#pragma unroll
                    for (int t = 0; t < 8; ++t) dst[t] = (TO)((double)a[t] + p.mu_const);
                }
            }
        }
    }
}

template <bool SOR, bool EMIT, typename TO>
__global__ void __launch_bounds__(256, 3) gibbs_torus_sweep_fused(const TorusSweepParams p, const int TR) {
    extern __shared__ __align__(128) float fsm[];
    const int h = p.h, n1 = p.n1, hv = h >> 2;
    const int r0 = blockIdx.x * TR;
    const int rows = min(TR, n1 - r0);
    const int64_t ch = blockIdx.y;
    const size_t plane = (size_t)n1 * h;
    const float* __restrict__ gred = p.red + ch * plane;      // read from the current buffers,
    const float* __restrict__ gblk = p.black + ch * plane;
    float* __restrict__ ored = p.red_out + ch * plane;        // write to the other pair: tiles of the
    float* __restrict__ oblk = p.black_out + ch * plane;      // same launch never see each other's output
    float* sB = fsm;                                       // [TR + 4][h]: black rows r0-2 ..
    float* sR = fsm + (size_t)(TR + 4) * h;                // [TR + 2][h]: red rows r0-1 ..
    uint64_t* bar = reinterpret_cast<uint64_t*>(fsm + (size_t)(2 * TR + 6) * h);
    const uint32_t row_bytes = (uint32_t)h * 4u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, row_bytes * (uint32_t)(2 * rows + 6));
        // rows are contiguous in HBM except at the torus wrap: at most three bulk copies per colour
        auto copy_rows = [&](float* dst, const float* src, int first, int cnt) {
            int done = 0;
            while (done < cnt) {
                int i = first + done;
                i = i < 0 ? i + n1 : (i >= n1 ? i - n1 : i);
                const int run = min(cnt - done, n1 - i);
                bulk_g2s(dst + (size_t)done * h, src + (size_t)i * h, row_bytes * (uint32_t)run, bar);
                done += run;
            }
        };
        copy_rows(sB, gblk, r0 - 2, rows + 4);
        copy_rows(sR, gred, r0 - 1, rows + 2);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    const uint64_t gch = (uint64_t)(p.chain0 + ch * p.chain_stride);
    const uint32_t c2 = (uint32_t)gch, c3 = (uint32_t)(gch >> 32) ^ (kTagTorus << 16);
    // warp ty owns rows ty, ty + nw, ...; its lanes stride the column vectors of the row
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, nw = blockDim.x >> 5;

    // ---- phase 1: red rows r0-1 .. r0+rows (local 0 .. rows+1), black neighbours from sB
    for (int lr = ty; lr < rows + 2; lr += nw) {
        int i = r0 - 1 + lr;
        i = i < 0 ? i + n1 : (i >= n1 ? i - n1 : i);
        const float* bo = sB + (size_t)(lr + 1) * h;
        float* rr = sR + (size_t)lr * h;
        float* gout = (lr >= 1 && lr <= rows) ? ored + (size_t)i * h : nullptr;
        const uint32_t ctr0 = (uint32_t)(i * hv);
        if (i & 1) fused_row<SOR, true, false, false, TO>(p, bo, rr, gout, nullptr, nullptr, false, tx, hv, h, ctr0, p.sweep * 2u, c2, c3);
        else fused_row<SOR, false, false, false, TO>(p, bo, rr, gout, nullptr, nullptr, false, tx, hv, h, ctr0, p.sweep * 2u, c2, c3);
    }
    __syncthreads();
    // ---- phase 2: black rows r0 .. r0+rows-1 from the new red rows in sR
    for (int t = ty; t < rows; t += nw) {
        const int i = r0 + t;
        const float* ro = sR + (size_t)(t + 1) * h;
        float* bc = sB + (size_t)(t + 2) * h;
        float* gout = oblk + (size_t)i * h;
        TO* erow = nullptr;
        const double* mrow = nullptr;
        if (EMIT) {
            if (ch < p.nch) {
                erow = reinterpret_cast<TO*>(p.out) + (size_t)(p.slot0 + ch) * p.ld + (size_t)i * 2 * h;
                if (p.mu != nullptr) mrow = p.mu + (size_t)i * 2 * h;
            }
        }
        const uint32_t ctr0 = (uint32_t)(i * hv);
        const bool par = (i & 1) != 0;
        if (!par) fused_row<SOR, true, true, EMIT, TO>(p, ro, bc, gout, erow, mrow, par, tx, hv, h, ctr0, p.sweep * 2u + 1u, c2, c3);
        else fused_row<SOR, false, true, EMIT, TO>(p, ro, bc, gout, erow, mrow, par, tx, hv, h, ctr0, p.sweep * 2u + 1u, c2, c3);
    }
}

// generic path: any h, one site per thread
template <bool SOR>
__global__ void __launch_bounds__(256) gibbs_torus_half_sweep_v1(const TorusSweepParams p, const int col) {
    const int h = p.h, n1 = p.n1;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    const int64_t ch = blockIdx.z;
    if (c >= h || i >= n1) return;
    const size_t plane = (size_t)n1 * h;
    float* own = (col == 0 ? p.red : p.black) + ch * plane;
    const float* oth = (col == 0 ? p.black : p.red) + ch * plane;
    const uint64_t gch = (uint64_t)(p.chain0 + ch * p.chain_stride);
    const int iu = i == 0 ? n1 - 1 : i - 1, id = i + 1 == n1 ? 0 : i + 1;
    const bool fwd = ((i & 1) ^ col) != 0;
    const int cn = fwd ? (c + 1 == h ? 0 : c + 1) : (c == 0 ? h - 1 : c - 1);
    const Philox4 r = philox4x32_10((uint32_t)(i * h + c), p.sweep * 2u + (uint32_t)col, (uint32_t)gch,
                                    (uint32_t)(gch >> 32) ^ (kTagTorus << 16), p.keys);
    float z0, z1;
    box_muller(r.x, r.y, z0, z1);
    const float s = oth[(size_t)i * h + c] + oth[(size_t)i * h + cn] + oth[(size_t)iu * h + c] + oth[(size_t)id * h + c];
    float v = fmaf(p.gain, s, p.noise * z0);
    if (SOR) v = fmaf(p.keep, own[(size_t)i * h + c], v);
    own[(size_t)i * h + c] = v;
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

template <int R>
void launch_v4(const TorusSweepParams& p, bool sor, bool emit, bool f32, int64_t C, cudaStream_t st) {
    const int hv = p.h / 4;
    const int strips = (p.n1 + R - 1) / R;
    dim3 block(256);  // 32 column vectors x 8 row strips
    dim3 grid((hv + 31) / 32, (strips + 7) / 8, (unsigned)C);
    if (sor) gibbs_torus_half_sweep_v4<true, false, float, R><<<grid, block, 0, st>>>(p, 0);
    else gibbs_torus_half_sweep_v4<false, false, float, R><<<grid, block, 0, st>>>(p, 0);
    if (emit) {
        if (sor && f32) gibbs_torus_half_sweep_v4<true, true, float, R><<<grid, block, 0, st>>>(p, 1);
        else if (sor) gibbs_torus_half_sweep_v4<true, true, double, R><<<grid, block, 0, st>>>(p, 1);
        else if (f32) gibbs_torus_half_sweep_v4<false, true, float, R><<<grid, block, 0, st>>>(p, 1);
        else gibbs_torus_half_sweep_v4<false, true, double, R><<<grid, block, 0, st>>>(p, 1);
    } else {
        if (sor) gibbs_torus_half_sweep_v4<true, false, float, R><<<grid, block, 0, st>>>(p, 1);
        else gibbs_torus_half_sweep_v4<false, false, float, R><<<grid, block, 0, st>>>(p, 1);
    }
}

}  // namespace

int torus_chains_alloc(mclcar_ctx* ctx, int64_t C, TorusChains& t) {
    const size_t bytes = (size_t)C * ctx->n1 * (ctx->n2 / 2) * sizeof(float);
    MCL_TRY(pool_alloc(ctx, (void**)&t.red, bytes));
    MCL_TRY(pool_alloc(ctx, (void**)&t.black, bytes));
    MCL_TRY(pool_alloc(ctx, (void**)&t.red2, bytes));
    MCL_TRY(pool_alloc(ctx, (void**)&t.black2, bytes));
    MCL_CUDA(ctx, cudaMemsetAsync(t.red, 0, bytes, ctx->stream));
    MCL_CUDA(ctx, cudaMemsetAsync(t.black, 0, bytes, ctx->stream));
    t.C = C;
    return MCLCAR_OK;
}

void torus_chains_free(mclcar_ctx* ctx, TorusChains& t) {
    pool_free(ctx, t.red);
    pool_free(ctx, t.black);
    pool_free(ctx, t.red2);
    pool_free(ctx, t.black2);
    t.red = t.black = t.red2 = t.black2 = nullptr;
}

// one full sweep (both colours) of all chains; omega in [1, 2): over-relaxation factor.  When
// `emit` is given (16-byte path only) the black half sweep also writes the samples.
int torus_sweep(mclcar_ctx* ctx, TorusChains& t, double rho, double sigma, double omega, uint32_t sweep,
                const TorusEmit* emit) {
    TorusSweepParams p{};
    p.red = t.red; p.black = t.black; p.n1 = ctx->n1; p.h = ctx->n2 / 2; p.C = t.C;
    p.keep = (float)(1.0 - omega);
    p.gain = (float)(omega * rho);
    p.noise = (float)std::sqrt(omega * (2.0 - omega) * sigma);
    p.rad2 = (float)(-2.0 * std::log(2.0) * omega * (2.0 - omega) * sigma);
    p.keys = philox_keys(ctx->draw_key);
    p.sweep = sweep;
    p.chain0 = ctx->rank; p.chain_stride = ctx->world;
    const bool vec4 = (p.h % 4) == 0;
    const bool sor = omega != 1.0;
    if (t.C > 65535) return fail(ctx, MCLCAR_ELIMIT, "more than 65535 chains in one torus sampler launch");
    if ((int64_t)p.n1 * p.h >= ((int64_t)1 << 31)) return fail(ctx, MCLCAR_ELIMIT, "lattice too large for 32-bit indexing");
    if (emit && !vec4) return fail(ctx, MCLCAR_EINVAL, "fused emission needs n2 % 8 == 0");
    if (vec4) {
        if (emit) {
            p.out = emit->out; p.ld = emit->ld; p.slot0 = emit->slot0; p.nch = emit->nch;
            p.mu = emit->d_mu; p.mu_const = emit->mu_const;
        }
        const bool f32 = !emit || emit->dtype == MCLCAR_F32;
        const char* fe = getenv("MCLCAR_SWEEP_FUSED");  // read per call: tests toggle it (0 = two half-sweep launches)
        const int fused_cfg = fe ? atoi(fe) : 1;
        static int tr_cfg = [] { const char* e = getenv("MCLCAR_SWEEP_TR"); return e ? atoi(e) : 0; }();
        const size_t row_b = (size_t)p.h * 4;
        const int fthreads = 256, nw = fthreads / 32;
        // tile height: three CTAs of <= 72 KB per SM; the red phase covers TR + 2 rows, so TR + 2 a multiple
        // of the warp count keeps the warps of a CTA level at the phase barrier (14 rows for 1000 x 1000)
        int TR = (int)((72 * 1024 / row_b - 6) / 2);
        TR = std::min(TR, 64);
        if (TR + 2 > nw && (TR + 2) % nw != 0) TR -= (TR + 2) % nw;
        if (tr_cfg > 0) TR = tr_cfg;
        if (fused_cfg && TR >= 4 && p.n1 >= TR + 4) {
            ProfScope prof(ctx, PROF_SWEEP, 1);
            p.red_out = t.red2; p.black_out = t.black2;
            std::swap(t.red, t.red2);        // the output pair is the current state from now on
            std::swap(t.black, t.black2);
            const size_t smem = (size_t)(2 * TR + 6) * row_b + 64;
            dim3 grid((p.n1 + TR - 1) / TR, (unsigned)t.C);
#define LAUNCH_FUSED(S, E, TO_)                                                                                  \
    do {                                                                                                         \
        MCL_CUDA(ctx, cudaFuncSetAttribute(gibbs_torus_sweep_fused<S, E, TO_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gibbs_torus_sweep_fused<S, E, TO_><<<grid, fthreads, smem, ctx->stream>>>(p, TR);                            \
    } while (0)
            if (emit) {
                if (sor && f32) LAUNCH_FUSED(true, true, float);
                else if (sor) LAUNCH_FUSED(true, true, double);
                else if (f32) LAUNCH_FUSED(false, true, float);
                else LAUNCH_FUSED(false, true, double);
            } else {
                if (sor) LAUNCH_FUSED(true, false, float);
                else LAUNCH_FUSED(false, false, float);
            }
#undef LAUNCH_FUSED
            MCL_CUDA(ctx, cudaGetLastError());
            return MCLCAR_OK;
        }
        ProfScope prof(ctx, PROF_SWEEP, 2);
        launch_v4<4>(p, sor, emit != nullptr, f32, t.C, ctx->stream);
    } else {
        ProfScope prof(ctx, PROF_SWEEP, 2);
        dim3 block(32, 8);
        dim3 grid((p.h + 31) / 32, (p.n1 + 7) / 8, (unsigned)t.C);
        for (int col = 0; col < 2; ++col) {
            if (sor) gibbs_torus_half_sweep_v1<true><<<grid, block, 0, ctx->stream>>>(p, col);
            else gibbs_torus_half_sweep_v1<false><<<grid, block, 0, ctx->stream>>>(p, col);
        }
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

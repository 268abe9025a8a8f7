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
constexpr int kFusedMaxThreads = 256;   // launch bound of the fused sweep (three CTAs per SM)

struct TorusSweepParams {
    float* red;    // [C][n1][h]
    float* black;  // [C][n1][h]
    float* red_out;    // fused sweep only: the other buffer of the ping-pong pair
    float* black_out;
    int n1, h;     // h = n2 / 2
    int G;         // Philox blocks per row = ceil(hv / 2), hv = h / 4 column vectors: block g serves vectors g and g + G
    uint32_t Gmagic;  // floor(2^32 / G) + 1: t / G == umulhi(t, Gmagic) for every item index t of a tile (checked on the host)
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
    const double* mu;   // [n] or null; for fp32 samples every 8-byte slot holds the fp32 pair (fl32(mu), fl32(mu - hi))
    double mu_const;    // used when mu is null
    // statistics-only emission (EMIT == 2): per (chain, tile) partial sums (x'x, x'Wx / 2, sum x) of the sample
    // fl32(state + mu_const), [C][tiles][3]
    double* stat_partial;
    int dry;              // experiment switch: skip the random numbers (structure / memory ceiling of the sweep)
    int prefetch_ahead;   // > 0: every CTA prefetches into L2 the tile of the CTA this many places behind it in launch order
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---------------------------------------------------------------------------------------------
// Random numbers.  One Philox4x32-10 block, counter row * G + g, serves the two 16-byte column vectors g and
// g + G of a row of one colour (G = ceil(vectors per row / 2)): consecutive lanes then work on consecutive
// vectors (coalesced stores, conflict-free shared-memory loads) and still share a block between their two
// vectors.  Every 32-bit word of the block is one Box-Muller pair (cos, sin -> even, odd column) -- radius
// from its top 23 bits (as in round 1), angle from its low 9 bits placed in the mantissa (no int -> float
// conversion) and centred in their cell.  A uniform angle on an N-point grid integrates every
// trigonometric polynomial of degree < N exactly, so all joint moments of (r cos, r sin) of total order
// < 512 are those of two independent normals, and the characteristic function of any projection
// a z0 + b z1 differs from the Gaussian one by Bessel terms J_512(|t| r) -- below 1e-30 for
// |t| r < 400: far beyond anything a fp32 state or the quadratic-form statistics of this path resolve
// (tests/test_gpu_sampler_dcar.py checks marginal and cross moments up to order 8 of the generator itself).
// Eight normals per block instead of four: Philox4x32-10 (20 IMAD.WIDE at quarter rate) was half of the
// cycles of the update (profiles/r01_sweep_experiments.md).  The mapping depends on (seed, draw, chain,
// sweep, colour, row, column) only: not on the launch geometry, the kernel form or the number of GPUs.
// ---------------------------------------------------------------------------------------------
struct Pair {
    float r, s, c;   // radius with the noise scale folded in (sqrt(rad2 * lg2 u), rad2 = -2 ln2 noise^2), sin, cos
};

__device__ __forceinline__ Pair bm_pair(const uint32_t w, const float rad2) {
    Pair p;
#ifdef MCLCAR_SWEEP_EXPERIMENTS
    if (rad2 > 0.f) {   // experiment switch (MCLCAR_SWEEP_DRY=2): no transcendental work (rad2 is negative in real runs)
        p.r = rad2; p.s = __uint_as_float(w); p.c = 0.5f;
        return p;
    }
#endif
    p.r = sqrt_approx(rad2 * lg2_approx(u01(w)));
    // 1 + (k + 1/2) / 512, k = low 9 bits: mantissa bits 22..14 and the half-step bit 13, exponent of 1.0 -- the angle
    // 2 pi u runs over [2 pi, 4 pi): sin and cos do not care about the whole turn, and the subtraction is saved
    uint32_t ub;                                  // (w & 0x1ff) * 2^14 + 0x3f802000 as and + multiply-add: two instructions
    asm("mad.lo.u32 %0, %1, 16384, 1065361408;" : "=r"(ub) : "r"(w & 0x1ffu));
    const float u = __uint_as_float(ub);
    sincos_approx(6.2831853071795865f * u, p.s, p.c);
    return p;
}
__device__ __forceinline__ Pair block_pair(const Philox4& b, const int t, const float rad2) {
    return bm_pair(t == 0 ? b.x : (t == 1 ? b.y : (t == 2 ? b.z : b.w)), rad2);
}

// the two pairs of ONE column vector (half-columns 4 cv .. 4 cv + 3) of a row, for the kernels that walk the
// lattice vector by vector (two-launch form): same numbers, one Philox block per vector
__device__ __forceinline__ void vector_pairs(const uint32_t row, const int cv, const int G, const uint32_t c1, const uint32_t c2,
                                             const uint32_t c3, const PhiloxKeys& keys, const float rad2, Pair& PA, Pair& PB) {
    const bool second = cv >= G;
    const Philox4 b = philox4x32_10(row * (uint32_t)G + (uint32_t)(second ? cv - G : cv), c1, c2, c3, keys);
    if (second) { PA = bm_pair(b.z, rad2); PB = bm_pair(b.w, rad2); }
    else { PA = bm_pair(b.x, rad2); PB = bm_pair(b.y, rad2); }
}

// The update of four consecutive half-columns of one row, shared by all kernel forms so that they produce the
// same bits:   x' = keep * x + gain * (sum of the four neighbours) + noise * z,   z ~ N(0, 1).
// o/up/d: the other colour at the same half-columns in this row / the row above / below; e: the
// same-row neighbour beyond the vector (c + 4 when FWD, c - 1 otherwise).  The noise scale is folded
// into the Box-Muller radius and the product radius * cos/sin into the final FMA: 6 FP32 instructions per
// site besides the RNG.
template <bool SOR, bool FWD>
__device__ __forceinline__ float4 gibbs_update4(const float4 o, const float4 up, const float4 d, const float e,
                                                const float4 cur, const Pair PA, const Pair PB, const float keep,
                                                const float gain) {
    float4 s;
    if (FWD) { s.x = o.x + o.y; s.y = o.y + o.z; s.z = o.z + o.w; s.w = o.w + e; }
    else { s.x = e + o.x; s.y = o.x + o.y; s.z = o.y + o.z; s.w = o.z + o.w; }
    s.x = __fadd_rn(s.x, __fadd_rn(up.x, d.x));
    s.y = __fadd_rn(s.y, __fadd_rn(up.y, d.y));
    s.z = __fadd_rn(s.z, __fadd_rn(up.z, d.z));
    s.w = __fadd_rn(s.w, __fadd_rn(up.w, d.w));
    float4 v;
    if (SOR) {
        v.x = __fmaf_rn(PA.r, PA.c, __fmul_rn(keep, cur.x));
        v.y = __fmaf_rn(PA.r, PA.s, __fmul_rn(keep, cur.y));
        v.z = __fmaf_rn(PB.r, PB.c, __fmul_rn(keep, cur.z));
        v.w = __fmaf_rn(PB.r, PB.s, __fmul_rn(keep, cur.w));
    } else {
        v.x = __fmul_rn(PA.r, PA.c); v.y = __fmul_rn(PA.r, PA.s); v.z = __fmul_rn(PB.r, PB.c); v.w = __fmul_rn(PB.r, PB.s);
    }
    v.x = __fmaf_rn(gain, s.x, v.x);
    v.y = __fmaf_rn(gain, s.y, v.y);
    v.z = __fmaf_rn(gain, s.z, v.z);
    v.w = __fmaf_rn(gain, s.w, v.w);
    return v;
}

// 16-byte global store at base + 4 * elem_off: one IMAD.WIDE.U32 for the address (the compiler's own pointer arithmetic
// re-adds the chain's plane offset and the buffer base with carries for every store)
__device__ __forceinline__ void stg4_off(const float* base, const uint32_t elem_off, const float4 v) {
    uint64_t a;
    asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(a) : "r"(elem_off), "l"(base));
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// fp32 sample = state + per-site mean, the mean given as an fp32 pair (hi, lo) = (fl32(mu), fl32(mu - hi)): TwoSum of state and
// hi, then the low parts -- the fp64 sum rounded to fp32 to within 2^-24 relative of half an ulp, in full-rate FP32
// instructions.  (fp64 add + two conversions per element ran on the conversion unit at 16 lanes per clock and made an
// emitting sweep with a non-constant mean 2.5 x as long as a plain one.)
__device__ __forceinline__ float add_mean_f32(const float a, const float2 m) {
    const float s = a + m.x;
    const float bb = s - a;
    const float err = (a - (s - bb)) + (m.x - bb);
    return s + (err + m.y);
}

// finished sample values of 8 consecutive sites (4 half-columns of both colours) -> the sample matrix.
// v: the emitting (black) colour, o: the other colour at the same half-columns; the emitting colour sits at column
// 2c + 1 - par, the other at 2c + par (par = row & 1).
template <typename TO>
__device__ __forceinline__ void emit8(const TorusSweepParams& p, TO* __restrict__ dst, const double* __restrict__ mu8,
                                      const float4 v, const float4 o, const bool par) {
    float a[8];
    a[0] = par ? v.x : o.x; a[1] = par ? o.x : v.x; a[2] = par ? v.y : o.y; a[3] = par ? o.y : v.y;
    a[4] = par ? v.z : o.z; a[5] = par ? o.z : v.z; a[6] = par ? v.w : o.w; a[7] = par ? o.w : v.w;
    if (mu8 != nullptr && sizeof(TO) == 4) {     // fp32 samples: the mean array holds (hi, lo) fp32 pairs
        const float4* m4 = reinterpret_cast<const float4*>(mu8);
        float r[8];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float4 mm = __ldg(m4 + t);
            r[2 * t] = add_mean_f32(a[2 * t], make_float2(mm.x, mm.y));
            r[2 * t + 1] = add_mean_f32(a[2 * t + 1], make_float2(mm.z, mm.w));
        }
        float4* d4 = reinterpret_cast<float4*>(dst);
        d4[0] = make_float4(r[0], r[1], r[2], r[3]);
        d4[1] = make_float4(r[4], r[5], r[6], r[7]);
    } else if (mu8 != nullptr) {
#pragma unroll
        for (int t = 0; t < 8; ++t) dst[t] = (TO)((double)a[t] + __ldg(mu8 + t));
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

// the same for the fused sweep: element offset from the chain's sample (one IMAD.WIDE per store) and the row parity known
// at compile time (tiles start at even rows: row A of a pair is even, row B odd)
template <typename TO, bool PAR>
__device__ __forceinline__ void emit8_off(const TorusSweepParams& p, TO* __restrict__ sample, const uint32_t off,
                                          const double* __restrict__ mu, const float4 v, const float4 o) {
    float a[8];
    a[0] = PAR ? v.x : o.x; a[1] = PAR ? o.x : v.x; a[2] = PAR ? v.y : o.y; a[3] = PAR ? o.y : v.y;
    a[4] = PAR ? v.z : o.z; a[5] = PAR ? o.z : v.z; a[6] = PAR ? v.w : o.w; a[7] = PAR ? o.w : v.w;
    if (mu != nullptr && sizeof(TO) == 4) {      // fp32 samples: the mean array holds (hi, lo) fp32 pairs
        const float4* m4 = reinterpret_cast<const float4*>(mu + off);
        float r[8];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float4 mm = __ldg(m4 + t);
            r[2 * t] = add_mean_f32(a[2 * t], make_float2(mm.x, mm.y));
            r[2 * t + 1] = add_mean_f32(a[2 * t + 1], make_float2(mm.z, mm.w));
        }
        stg4_off(reinterpret_cast<const float*>(sample), off, make_float4(r[0], r[1], r[2], r[3]));
        stg4_off(reinterpret_cast<const float*>(sample), off + 4u, make_float4(r[4], r[5], r[6], r[7]));
    } else if (mu != nullptr) {
#pragma unroll
        for (int t = 0; t < 8; ++t) sample[off + t] = (TO)((double)a[t] + __ldg(mu + off + t));
    } else if (sizeof(TO) == 4) {
        const float m = (float)p.mu_const;
        stg4_off(reinterpret_cast<const float*>(sample), off, make_float4(a[0] + m, a[1] + m, a[2] + m, a[3] + m));
        stg4_off(reinterpret_cast<const float*>(sample), off + 4u, make_float4(a[4] + m, a[5] + m, a[6] + m, a[7] + m));
    } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) sample[off + t] = (TO)((double)a[t] + p.mu_const);
    }
}

// Statistics-only emission: the contribution of one black vector v (new values) to (x'x, x'Wx / 2, sum x) of the
// fp32 sample y = fl32(state + m) -- exactly what the statistics kernel would compute from the stored matrix.
// Every edge of the even torus joins a black and a red site, so x'Wx = 2 sum_black y_b * (sum of its four red
// neighbours); o/up/dn/e are those neighbours (the finished red values), o also contributes its own squares.
template <bool FWD>
__device__ __forceinline__ void stats8(const float4 v, const float4 o, const float4 up, const float4 dn, const float e,
                                       const float m, double (&acc)[3]) {
    const double dv[4] = {(double)(v.x + m), (double)(v.y + m), (double)(v.z + m), (double)(v.w + m)};
    const double dO[4] = {(double)(o.x + m), (double)(o.y + m), (double)(o.z + m), (double)(o.w + m)};
    const double du[4] = {(double)(up.x + m), (double)(up.y + m), (double)(up.z + m), (double)(up.w + m)};
    const double dd[4] = {(double)(dn.x + m), (double)(dn.y + m), (double)(dn.z + m), (double)(dn.w + m)};
    const double de = (double)(e + m);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double side = FWD ? (k < 3 ? dO[(k + 1) & 3] : de) : (k > 0 ? dO[(k + 3) & 3] : de);
        const double S = (dO[k] + side) + (du[k] + dd[k]);
        acc[1] = fma(dv[k], S, acc[1]);
        acc[0] = fma(dv[k], dv[k], acc[0]);
        acc[0] = fma(dO[k], dO[k], acc[0]);
        acc[2] += dv[k] + dO[k];
    }
}

// 16-byte path (h % 4 == 0), two launches per sweep.  SOR: the update reads the current value too.  EMIT: the
// black half sweep also stores the finished sample row segment (both colours are in registers).
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
        Pair PA, PB;
        vector_pairs((uint32_t)i, cv, p.G, c1, c2, c3, p.keys, p.rad2, PA, PB);
        const float4 v = fwd ? gibbs_update4<SOR, true>(o, up, d, e, cur, PA, PB, p.keep, p.gain)
                             : gibbs_update4<SOR, false>(o, up, d, e, cur, PA, PB, p.keep, p.gain);
        st4(own + row + c0, v);
        if (EMIT) {
            if (ch < p.nch) {
                // black half sweep (col == 1): o = finished red values, v = new black values of the same half-columns
                const size_t site = (size_t)i * 2 * h + 2 * c0;
                TO* dst = reinterpret_cast<TO*>(p.out) + (size_t)(p.slot0 + ch) * p.ld + site;
                emit8<TO>(p, dst, p.mu ? p.mu + site : nullptr, v, o, (i & 1) != 0);
            }
        }
        up = o;
        o = d;
        d = dn;
        e = en;
        cur = curn;
        row = rown;
        rown = row2;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused full sweep: ONE pass over HBM per sweep instead of two.  A CTA owns TR (even) lattice rows of
// one chain at full width.  Black rows r0-2 .. r0+TR+1 and red rows r0-1 .. r0+TR are staged into
// shared memory by the bulk-async copy engine (cp.async.bulk, one mbarrier); phase 1 updates red
// for the TR rows AND the two halo rows (recomputing what the neighbouring tiles compute -- same
// Philox counters, hence the same numbers), phase 2 updates black for the TR rows from the new
// red values in shared memory.  Traffic: (2TR+6)/(2TR) reads + 1 write of the state per sweep,
// about 9 B per site update instead of 12 (over-relaxed) for the two-launch form; results are
// bit-identical to it.
//
// Work item = (pair of vertically adjacent rows A, B = A + 1; Philox block g = column vectors g and g + G): 16 site
// updates from TWO Philox blocks.  The pair shares its neighbour-row loads (rows A-1, A, B, B+1 of the other
// colour: 8 vector loads for 4 vector updates instead of 12); consecutive lanes hold consecutive vectors, so
// the same-row neighbours come from the next / previous lane by warp shuffle.  The items of a phase are dealt to the
// threads of the CTA round-robin, so only the last round of a phase has idle lanes.  Tiles start at even
// rows, hence row A is always the row whose same-row neighbour sits at c + 1 (FWD) and row B the one with
// c - 1, in both phases: no divergent code paths.
// ---------------------------------------------------------------------------------------------
// Shared memory is addressed through explicit 32-bit shared-space addresses (ld.shared / st.shared): with generic
// pointers the compiler re-derives the shared window for every access (the kernel uses the cluster-scoped bulk copy,
// so the window depends on the CTA's place in its cluster) -- ~60 instructions per item in round 1's SASS.
__device__ __forceinline__ float4 lds4(const uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds1(const uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts4(const uint32_t a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// EMIT: 0 none, 1 the black phase also writes the finished samples, 2 it accumulates their statistics instead, 3 both.
// nm: shared address of row A of the OTHER colour (rows A-1, A, B, B+1 at nm - hb .. nm + 2 hb, hb = 4 h bytes);
// own: shared address of row A of the colour being updated; gout: the output half-lattice of the chain, offA / offB
// the element offsets of rows A and B in it (stA / stB: whether the row is stored -- halo rows are not).
template <bool SOR, bool PH2, int EMIT, typename TO, bool STS>
__device__ __forceinline__ void fused_item(const TorusSweepParams& p, const bool valid, const uint32_t nm, const uint32_t own,
                                           float* __restrict__ gout, const uint32_t offA, const uint32_t offB, const bool stA, const bool stB,
                                           TO* __restrict__ esample, const uint32_t eoff, const double* __restrict__ mu, const uint32_t iA,
                                           const uint32_t iB, const int g, const int hv, const int h, const uint32_t c1,
                                           const uint32_t c2, const uint32_t c3, double (&acc)[3]) {
    const int G = p.G;
    const int gb = g + G;
    const bool two = valid && gb < hv;                        // the last block of a row may hold one vector only
    const int c0 = 4 * g, c1v = 4 * gb;
    const uint32_t hb = (uint32_t)h * 4u, b0a = (uint32_t)c0 * 4u, b1a = (uint32_t)c1v * 4u;   // byte offsets
    if (!valid) return;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 a0 = lds4(nm + b0a), b0 = lds4(nm + hb + b0a);           // rows A and B of the other colour
    const float4 a1 = two ? lds4(nm + b1a) : z4, b1 = two ? lds4(nm + hb + b1a) : z4;
    // same-row neighbours beyond the vector: row A looks forward (c + 4, cyclic), row B backward (c - 1, cyclic)
    const float eA0 = lds1(nm + (g + 1 == hv ? 0u : b0a + 16u));
    const float eA1 = two ? lds1(nm + (gb + 1 == hv ? 0u : b1a + 16u)) : 0.f;
    const float eB0 = lds1(nm + hb + (g == 0 ? hb - 4u : b0a - 4u));
    const float eB1 = two ? lds1(nm + hb + b1a - 4u) : 0.f;
    const uint32_t ownB = own + hb;
    const float m = (float)p.mu_const;
    // ---- row A
    {
#ifdef MCLCAR_SWEEP_EXPERIMENTS
        const Philox4 blk = p.dry ? Philox4{0x80000000u + iA, 0xa0000000u + c1, 0x90000000u + (uint32_t)g, 0xc0000000u + c2}
                                  : philox4x32_10(iA * (uint32_t)G + (uint32_t)g, c1, c2, c3, p.keys);
#else
        const Philox4 blk = philox4x32_10(iA * (uint32_t)G + (uint32_t)g, c1, c2, c3, p.keys);
#endif
        {
            const float4 up = lds4(nm - hb + b0a);
            const float4 cur = SOR ? lds4(own + b0a) : z4;
            const float4 v = gibbs_update4<SOR, true>(a0, up, b0, eA0, cur, bm_pair(blk.x, p.rad2), bm_pair(blk.y, p.rad2), p.keep, p.gain);
            if (STS) sts4(own + b0a, v);
            if (stA) stg4_off(gout, offA + (uint32_t)c0, v);
            if (PH2 && (EMIT == 1 || EMIT == 3)) {
                if (esample) emit8_off<TO, false>(p, esample, eoff + 2u * (uint32_t)c0, mu, v, a0);
            }
            if (PH2 && EMIT >= 2) stats8<true>(v, a0, up, b0, eA0, m, acc);
        }
        if (two) {
            const float4 up = lds4(nm - hb + b1a);
            const float4 cur = SOR ? lds4(own + b1a) : z4;
            const float4 v = gibbs_update4<SOR, true>(a1, up, b1, eA1, cur, bm_pair(blk.z, p.rad2), bm_pair(blk.w, p.rad2), p.keep, p.gain);
            if (STS) sts4(own + b1a, v);
            if (stA) stg4_off(gout, offA + (uint32_t)c1v, v);
            if (PH2 && (EMIT == 1 || EMIT == 3)) {
                if (esample) emit8_off<TO, false>(p, esample, eoff + 2u * (uint32_t)c1v, mu, v, a1);
            }
            if (PH2 && EMIT >= 2) stats8<true>(v, a1, up, b1, eA1, m, acc);
        }
    }
    // ---- row B
    {
#ifdef MCLCAR_SWEEP_EXPERIMENTS
        const Philox4 blk = p.dry ? Philox4{0x80000000u + iB, 0xa0000000u + c1, 0x90000000u + (uint32_t)g, 0xc0000000u + c2}
                                  : philox4x32_10(iB * (uint32_t)G + (uint32_t)g, c1, c2, c3, p.keys);
#else
        const Philox4 blk = philox4x32_10(iB * (uint32_t)G + (uint32_t)g, c1, c2, c3, p.keys);
#endif
        {
            const float4 dn = lds4(nm + 2u * hb + b0a);
            const float4 cur = SOR ? lds4(ownB + b0a) : z4;
            const float4 v = gibbs_update4<SOR, false>(b0, a0, dn, eB0, cur, bm_pair(blk.x, p.rad2), bm_pair(blk.y, p.rad2), p.keep, p.gain);
            if (STS) sts4(ownB + b0a, v);
            if (stB) stg4_off(gout, offB + (uint32_t)c0, v);
            if (PH2 && (EMIT == 1 || EMIT == 3)) {
                if (esample) emit8_off<TO, true>(p, esample, eoff + 2u * (uint32_t)(h + c0), mu, v, b0);
            }
            if (PH2 && EMIT >= 2) stats8<false>(v, b0, a0, dn, eB0, m, acc);
        }
        if (two) {
            const float4 dn = lds4(nm + 2u * hb + b1a);
            const float4 cur = SOR ? lds4(ownB + b1a) : z4;
            const float4 v = gibbs_update4<SOR, false>(b1, a1, dn, eB1, cur, bm_pair(blk.z, p.rad2), bm_pair(blk.w, p.rad2), p.keep, p.gain);
            if (STS) sts4(ownB + b1a, v);
            if (stB) stg4_off(gout, offB + (uint32_t)c1v, v);
            if (PH2 && (EMIT == 1 || EMIT == 3)) {
                if (esample) emit8_off<TO, true>(p, esample, eoff + 2u * (uint32_t)(h + c1v), mu, v, b1);
            }
            if (PH2 && EMIT >= 2) stats8<false>(v, b1, a1, dn, eB1, m, acc);
        }
    }
}

// NS: sweeps per launch.  NS = 2 keeps a tile in shared memory for TWO full sweeps (red, black, red, black; the region
// that can be advanced shrinks by one row per half sweep on either side, so the tile is loaded with 4 NS - 1 halo rows
// per colour end and the halo is recomputed from the same Philox counters as the neighbouring tiles): the state crosses
// HBM once per two sweeps.  Only the last sweep's rows go back to global memory; the results are those of NS single launches,
// bit for bit.  Emission (EMIT != 0) belongs to single-sweep launches.
template <bool SOR, int EMIT, typename TO, int MINB, int NS, int MAXT>
__global__ void __launch_bounds__(MAXT, MINB) gibbs_torus_sweep_fused(const TorusSweepParams p, const int TR) {
    extern __shared__ __align__(128) float fsm[];
    constexpr int HB = 2 * NS, HR = 2 * NS - 1;             // halo rows above the tile: black, red
    const int h = p.h, n1 = p.n1, hv = h >> 2, G = p.G;
    const int r0 = blockIdx.x * TR;
    const int rows = min(TR, n1 - r0);                        // even: TR and n1 are
    const int64_t ch = blockIdx.y;
    const size_t plane = (size_t)n1 * h;
    const float* __restrict__ gred = p.red + ch * plane;      // read from the current buffers,
    const float* __restrict__ gblk = p.black + ch * plane;
    float* __restrict__ ored = p.red_out + ch * plane;        // write to the other pair: tiles of the
    float* __restrict__ oblk = p.black_out + ch * plane;      // same launch never see each other's output
    float* sB = fsm;                                       // [TR + 2 HB][h]: black rows r0 - HB ..
    float* sR = fsm + (size_t)(TR + 2 * HB) * h;           // [TR + 2 HR][h]: red rows r0 - HR ..
    uint64_t* bar = reinterpret_cast<uint64_t*>(fsm + (size_t)(2 * TR + 2 * HB + 2 * HR) * h);
    const uint32_t row_bytes = (uint32_t)h * 4u;
    const uint32_t aB = smem_u32(sB), aR = smem_u32(sR);   // shared-space addresses
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, row_bytes * (uint32_t)(2 * rows + 2 * HB + 2 * HR));
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
        copy_rows(sB, gblk, r0 - HB, rows + 2 * HB);
        copy_rows(sR, gred, r0 - HR, rows + 2 * HR);
#ifdef MCLCAR_SWEEP_EXPERIMENTS
        if (p.prefetch_ahead > 0) {
            // warm L2 with the tile of the CTA that will follow this one onto the SM: CTAs are dispatched in linear
            // order, about (resident CTAs of the device) behind the running front
            const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + p.prefetch_ahead;
            const long long pch = lin / gridDim.x;
            if (pch < (long long)gridDim.y) {
                const int pr0 = (int)(lin - pch * gridDim.x) * TR;
                const int prow = min(TR, n1 - pr0);
                const float* pb = p.black + pch * plane + (size_t)pr0 * h;
                const float* pr = p.red + pch * plane + (size_t)pr0 * h;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pb), "r"(row_bytes * (uint32_t)prow) : "memory");
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pr), "r"(row_bytes * (uint32_t)prow) : "memory");
            }
        }
#endif
    }
    __syncthreads();
    mbar_wait(bar, 0);

    const uint64_t gch = (uint64_t)(p.chain0 + ch * p.chain_stride);
    const uint32_t c2 = (uint32_t)gch, c3 = (uint32_t)(gch >> 32) ^ (kTagTorus << 16);
    double acc[3] = {0.0, 0.0, 0.0};
    TO* esample = nullptr;                                    // this chain's sample in the output matrix (emitting sweeps)
    if (EMIT == 1 || EMIT == 3) {
        if (ch < p.nch) esample = reinterpret_cast<TO*>(p.out) + (size_t)(p.slot0 + ch) * p.ld;
    }

#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const bool last = k == NS - 1;
        const uint32_t swp = (p.sweep + (uint32_t)k) * 2u;
        if (k > 0) __syncthreads();
        // ---- red rows r0 - HR + 2k .. r0 + rows + HR - 2k - 1 (local 2k ..) in pairs, black neighbours from sB
        {
            const int items = ((rows + 2 * HR - 4 * k) >> 1) * G;
            for (int t0 = 0; t0 < items; t0 += blockDim.x) {     // uniform trip count
                const int t = t0 + threadIdx.x;
                const bool valid = t < items;
                const int q = valid ? (int)__umulhi((uint32_t)t, p.Gmagic) : 0, g = valid ? t - q * G : 0;
                const int lA = 2 * k + 2 * q;                      // local red row of A in sR; its black row in sB is lA + 1
                int iA = r0 - HR + lA;
                iA = iA < 0 ? iA + n1 : (iA >= n1 ? iA - n1 : iA);
                int iB = r0 - HR + lA + 1;
                iB = iB < 0 ? iB + n1 : (iB >= n1 ? iB - n1 : iB);
                // only the last sweep's rows inside the tile are stored (halo rows are recomputed by the neighbours)
                const bool stA = last && lA >= HR && lA < HR + rows, stB = last && lA + 1 >= HR && lA + 1 < HR + rows;
                fused_item<SOR, false, 0, TO, true>(p, valid, aB + (uint32_t)(lA + 1) * row_bytes, aR + (uint32_t)lA * row_bytes, ored,
                                                    (uint32_t)(iA * h), (uint32_t)(iB * h), stA, stB, nullptr, 0u, nullptr, (uint32_t)iA,
                                                    (uint32_t)iB, g, hv, h, swp, c2, c3, acc);
            }
        }
        __syncthreads();
        // ---- black rows r0 - HR + 2k + 1 .. (local 2k + 2 .. in sB) from the new red rows in sR
        {
            const int items = ((rows + 2 * HR - 2 - 4 * k) >> 1) * G;
            for (int t0 = 0; t0 < items; t0 += blockDim.x) {
                const int t = t0 + threadIdx.x;
                const bool valid = t < items;
                const int q = valid ? (int)__umulhi((uint32_t)t, p.Gmagic) : 0, g = valid ? t - q * G : 0;
                const int lb = 2 * k + 2 + 2 * q;                  // local black row of A in sB; its red row in sR is lb - 1
                int iA = r0 - HB + lb;
                iA = iA < 0 ? iA + n1 : (iA >= n1 ? iA - n1 : iA);
                int iB = r0 - HB + lb + 1;
                iB = iB < 0 ? iB + n1 : (iB >= n1 ? iB - n1 : iB);
                if (last)      // exactly the tile's rows: stored (and emitted)
                    fused_item<SOR, true, EMIT, TO, false>(p, valid, aR + (uint32_t)(lb - 1) * row_bytes, aB + (uint32_t)lb * row_bytes, oblk,
                                                           (uint32_t)(iA * h), (uint32_t)(iB * h), true, true, esample, (uint32_t)(iA * 2 * h), p.mu,
                                                           (uint32_t)iA, (uint32_t)iB, g, hv, h, swp + 1u, c2, c3, acc);
                else           // an inner sweep: the new black rows stay in shared memory for the next red phase
                    fused_item<SOR, true, 0, TO, true>(p, valid, aR + (uint32_t)(lb - 1) * row_bytes, aB + (uint32_t)lb * row_bytes, oblk,
                                                       (uint32_t)(iA * h), (uint32_t)(iB * h), false, false, nullptr, 0u, nullptr,
                                                       (uint32_t)iA, (uint32_t)iB, g, hv, h, swp + 1u, c2, c3, acc);
            }
        }
    }
    if (EMIT >= 2) {
        double* scratch = reinterpret_cast<double*>(bar + 2);
        block_sum<3>(acc, scratch);
        if (threadIdx.x == 0) {
            double* o = p.stat_partial + ((size_t)ch * gridDim.x + blockIdx.x) * 3;
            o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
        }
    }
}

// statistics-only emission, second step: add the tile partials of every chain in tile order and expand them to
// the d statistic rows: x'x, x'Wx, and for the (constant) design columns X'x = c sum x, (WX)'x = 4 c sum x
struct StatFinishParams {
    const double* partial;   // [C][tiles][3]
    int tiles, d, nconst;
    int dst_a[MCLCAR_MAX_COV], dst_b[MCLCAR_MAX_COV];
    double cval[MCLCAR_MAX_COV];
    double* q;               // [d][ldq], first sample of this emission at q[slot0]
    int64_t ldq, slot0, nch;
};
__global__ void torus_stats_finish_kernel(const StatFinishParams f) {
    const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= f.nch) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const double* pp = f.partial + (size_t)ch * f.tiles * 3;
    for (int t = 0; t < f.tiles; ++t) { a0 += pp[3 * t]; a1 += pp[3 * t + 1]; a2 += pp[3 * t + 2]; }
    double* o = f.q + f.slot0 + ch;
    o[0] = a0;
    o[f.ldq] = 2.0 * a1;
    for (int l = 0; l < f.nconst; ++l) {
        o[(size_t)f.dst_a[l] * f.ldq] = f.cval[l] * a2;
        o[(size_t)f.dst_b[l] * f.ldq] = 4.0 * f.cval[l] * a2;
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
                                                         const double* mu, double mu_const, TO* out, int64_t ld,
                                                         int64_t slot0, int64_t nch) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    const int64_t ch = blockIdx.z;
    if (c >= h || ch >= nch) return;
    const size_t plane = (size_t)n1 * h;
    const float r = red[ch * plane + (size_t)i * h + c];
    const float b = black[ch * plane + (size_t)i * h + c];
    const int par = i & 1;  // red sits at column 2c + par, black at 2c + 1 - par
    const size_t site = (size_t)i * 2 * h + 2 * c;
    TO* o = out + (size_t)(slot0 + ch) * ld + site;
    if (mu != nullptr && sizeof(TO) == 4) {      // (hi, lo) fp32 pairs, as in emit8
        const float4 mm = *reinterpret_cast<const float4*>(mu + site);
        o[0] = (TO)add_mean_f32(par ? b : r, make_float2(mm.x, mm.y));
        o[1] = (TO)add_mean_f32(par ? r : b, make_float2(mm.z, mm.w));
        return;
    }
    if (mu == nullptr && sizeof(TO) == 4) {      // one mean for every site: fp32 sum, as in emit8
        const float m = (float)mu_const;
        o[0] = (TO)((par ? b : r) + m);
        o[1] = (TO)((par ? r : b) + m);
        return;
    }
    const double m0 = mu ? mu[site] : mu_const, m1 = mu ? mu[site + 1] : mu_const;
    const double v0 = (par ? (double)b : (double)r) + m0;
    const double v1 = (par ? (double)r : (double)b) + m1;
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

// tile height of the fused sweep for this context's lattice (0: the fused sweep does not apply).  Three CTAs of <= 72 KB
// per SM (a 4-CTA build at 64 registers and 10-row tiles measured 17 % slower: profiles/r02_sweep_experiments.md); even,
// so that every tile starts at an even row (row-pair work items)
int torus_fused_tile_rows(const mclcar_ctx* ctx) {
    const int h = ctx->n2 / 2;
    if (ctx->n1 % 2 != 0 || ctx->n2 % 2 != 0 || h % 4 != 0) return 0;
    const char* fe = getenv("MCLCAR_SWEEP_FUSED");  // read per call: tests toggle it (0 = two half-sweep launches)
    if (fe && atoi(fe) == 0) return 0;
    const size_t row_b = (size_t)h * 4;
    int TR = (int)((72 * 1024 / row_b - 6) / 2);
    TR = std::min(TR, 64) & ~1;
    if (const char* e = getenv("MCLCAR_SWEEP_TR")) {
        const int v = atoi(e);
        if (v > 0) TR = v & ~1;
    }
    if (TR < 4 || ctx->n1 < TR + 4) return 0;
    return TR;
}

// one full sweep (both colours) of all chains; omega in [1, 2): over-relaxation factor.  When
// `emit` is given (16-byte path only) the black half sweep also writes the samples; with emit->stats_q set
// (fused sweep, constant mean, fp32 samples) it forms their sufficient statistics -- instead of the samples
// (emit->out null) or together with them.
// tile height of the two-sweeps-per-launch form (0: not available): two CTAs of <= 110 KB per SM, 4 + 3 halo rows per
// colour and tile end
// Measured (profiles/r02_sweep_experiments.md): under the power cap of a sustained full-size run the two-sweep form is
// 7 % faster (less energy per site update: 4.7 instead of 8 bytes of HBM traffic, SM clock 1.93 instead of 1.79 GHz);
// in short runs at boost clocks it is 3 % slower (two larger CTAs per SM hide latency less well than three).  It is
// therefore used for launches of >= 10^8 site updates -- the sustained regime -- unless MCLCAR_SWEEP_DOUBLE = 1 | 0 says otherwise.
int torus_fused2_tile_rows(const mclcar_ctx* ctx, int64_t chains) {
    if (!torus_fused_tile_rows(ctx)) return 0;
    const char* e2 = getenv("MCLCAR_SWEEP_DOUBLE");
    if (e2 && atoi(e2) == 0) return 0;
    if (!(e2 && atoi(e2) != 0) && chains * (int64_t)ctx->n < 100000000ll) return 0;
    const size_t row_b = (size_t)(ctx->n2 / 2) * 4;
    int TR = (int)((110 * 1024 / row_b - 14) / 2);
    TR = std::min(TR, 64) & ~1;
    if (const char* e = getenv("MCLCAR_SWEEP_TR2")) {
        const int v = atoi(e);
        if (v > 0) TR = v & ~1;
    }
    if (TR < 8 || ctx->n1 < TR + 8) return 0;
    return TR;
}

// Tile height of the statistics-forming emitting sweeps.  They run two CTAs per SM (126 registers), so a tile may use
// up to ~110 KB: taller tiles recompute fewer halo rows and pay the tile load and its barriers less often -- but a
// phase takes ceil(items / threads) rounds of the CTA, and a height whose last round is nearly empty wastes it.  The
// height is the one with the fewest rounds per sweep of a chain, one round charged per tile for its load; measured on
// the 1000 x 1000 lattice (full-size bench, 84 emitting launches per step): 14 rows 0.363 ms per launch, 18 rows 0.355,
// **22 rows 0.323** (the model's choice), 26 rows do not fit.  MCLCAR_EMIT_TR overrides.
int torus_stats_tile_rows(const mclcar_ctx* ctx, int tr_default) {
    const int h = ctx->n2 / 2, G = (h / 4 + 1) / 2, n1 = ctx->n1, threads = 256;
    auto fits = [&](int tr) { return tr >= 4 && n1 >= tr + 4 && (size_t)(2 * tr + 6) * h * 4 + 1024 <= 110 * 1024; };
    if (const char* e = getenv("MCLCAR_EMIT_TR")) {
        const int v = atoi(e) & ~1;
        return fits(v) ? v : tr_default;
    }
    auto rounds = [&](int rows) {   // red phase: rows + 2 rows in pairs; black phase: rows
        return ((rows + 2) / 2 * G + threads - 1) / threads + (rows / 2 * G + threads - 1) / threads + 1;
    };
    int best = tr_default;
    long best_cost = -1;
    for (int tr = 64; tr >= 8; tr -= 2) {
        if (!fits(tr)) continue;
        const int full = n1 / tr, rest = n1 - full * tr;
        const long cost = (long)full * rounds(tr) + (rest ? rounds(rest) : 0);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = tr; }
    }
    return best;
}

// `want` >= 2 without emission: two sweeps (sweep, sweep + 1) in one launch where the lattice allows; *done = sweeps made
int torus_sweep(mclcar_ctx* ctx, TorusChains& t, double rho, double sigma, double omega, uint32_t sweep,
                const TorusEmit* emit, int want, int* done) {
    if (done) *done = 1;
    TorusSweepParams p{};
    p.red = t.red; p.black = t.black; p.n1 = ctx->n1; p.h = ctx->n2 / 2; p.C = t.C;
    p.G = (p.h / 4 + 1) / 2;
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
        const bool want_stats = emit && emit->stats_q != nullptr;
        const int TR2 = (want >= 2 && !emit && done) ? torus_fused2_tile_rows(ctx, t.C) : 0;
        int TR = TR2 ? TR2 : torus_fused_tile_rows(ctx);
        if (TR && !TR2 && want_stats) TR = torus_stats_tile_rows(ctx, TR);
        if (want_stats && (!TR || emit->d_mu || !f32))
            return fail(ctx, MCLCAR_ESTATE, "statistics-only emission needs the fused sweep, a constant mean and fp32 samples");
        if (TR) {
            const size_t row_b = (size_t)p.h * 4;
            const int tiles = (p.n1 + TR - 1) / TR;
            if (want_stats) {
                MCL_TRY(ensure(ctx, ctx->d_partial, (size_t)t.C * tiles * 3 * sizeof(double)));
                p.stat_partial = (double*)ctx->d_partial.p;
            }
            ProfScope prof(ctx, emit ? PROF_EMIT : PROF_SWEEP, 1);   // emitting launches are timed apart: they also write the samples
#ifdef MCLCAR_SWEEP_EXPERIMENTS   // build with -DMCLCAR_SWEEP_EXPERIMENTS for the switches of profiles/r02_sweep_experiments.md
            if (const char* e = getenv("MCLCAR_SWEEP_PREFETCH")) p.prefetch_ahead = atoi(e);
            if (const char* e = getenv("MCLCAR_SWEEP_DRY")) { p.dry = atoi(e); if (p.dry >= 2) p.rad2 = 1e-30f; }
#endif
            p.red_out = t.red2; p.black_out = t.black2;
            std::swap(t.red, t.red2);        // the output pair is the current state from now on
            std::swap(t.black, t.black2);
            {   // item index -> (row pair, block) by a multiply: exact for every index a tile can have
                p.Gmagic = (uint32_t)(((uint64_t)1 << 32) / (uint64_t)p.G) + 1u;
                const uint32_t tmax = (uint32_t)(((TR + 6) / 2) * p.G + 2 * kFusedMaxThreads);
                if ((uint64_t)tmax * p.G >= ((uint64_t)1 << 32)) return fail(ctx, MCLCAR_ELIMIT, "lattice row too long for the fused sweep");
            }
            int fthreads = 256;
            if (const char* e = getenv("MCLCAR_SWEEP_THREADS")) fthreads = std::min(std::max(atoi(e), 64), kFusedMaxThreads) & ~31;
            size_t smem = (size_t)(2 * TR + 6) * row_b + 64 + (kFusedMaxThreads / 32) * 3 * sizeof(double);
            dim3 grid(tiles, (unsigned)t.C);
            if (TR2) {      // two sweeps in one launch
                constexpr int kT2 = 448;
                int th2 = kT2;
                if (const char* e = getenv("MCLCAR_SWEEP2_THREADS")) th2 = std::min(std::max(atoi(e), 64), kT2) & ~31;
                smem = (size_t)(2 * TR + 14) * row_b + 64;
                if (sor) {
                    MCL_CUDA(ctx, cudaFuncSetAttribute(gibbs_torus_sweep_fused<true, 0, float, 2, 2, kT2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    gibbs_torus_sweep_fused<true, 0, float, 2, 2, kT2><<<grid, th2, smem, ctx->stream>>>(p, TR);
                } else {
                    MCL_CUDA(ctx, cudaFuncSetAttribute(gibbs_torus_sweep_fused<false, 0, float, 2, 2, kT2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    gibbs_torus_sweep_fused<false, 0, float, 2, 2, kT2><<<grid, th2, smem, ctx->stream>>>(p, TR);
                }
                MCL_CUDA(ctx, cudaGetLastError());
                *done = 2;
                return MCLCAR_OK;
            }
#define LAUNCH_FUSED_B(S, E, TO_, B_)                                                                            \
    do {                                                                                                         \
        MCL_CUDA(ctx, cudaFuncSetAttribute(gibbs_torus_sweep_fused<S, E, TO_, B_, 1, kFusedMaxThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gibbs_torus_sweep_fused<S, E, TO_, B_, 1, kFusedMaxThreads><<<grid, fthreads, smem, ctx->stream>>>(p, TR);   \
    } while (0)
#define LAUNCH_FUSED(S, E, TO_) LAUNCH_FUSED_B(S, E, TO_, 3)
            if (want_stats && emit->out) {                       // samples and their statistics from the same registers
                // two CTAs per SM (126 registers); at three (80 registers) the fp64 accumulation spills and the sweep is 9 % slower
                if (sor) LAUNCH_FUSED_B(true, 3, float, 2);
                else LAUNCH_FUSED_B(false, 3, float, 2);
            } else if (want_stats) {
                if (sor) LAUNCH_FUSED_B(true, 2, float, 2);      // fp64 accumulators: two CTAs per SM
                else LAUNCH_FUSED_B(false, 2, float, 2);
            } else if (emit) {
                if (sor && f32) LAUNCH_FUSED(true, 1, float);
                else if (sor) LAUNCH_FUSED(true, 1, double);
                else if (f32) LAUNCH_FUSED(false, 1, float);
                else LAUNCH_FUSED(false, 1, double);
            } else {
                if (sor) LAUNCH_FUSED(true, 0, float);
                else LAUNCH_FUSED(false, 0, float);
            }
#undef LAUNCH_FUSED
#undef LAUNCH_FUSED_B
            MCL_CUDA(ctx, cudaGetLastError());
            if (want_stats) {
                StatFinishParams f{};
                f.partial = p.stat_partial; f.tiles = tiles; f.d = 2 + 2 * ctx->p; f.nconst = ctx->p;
                for (int l = 0; l < ctx->p; ++l) {
                    f.dst_a[l] = 2 + l; f.dst_b[l] = 2 + ctx->p + l; f.cval[l] = ctx->X[(size_t)l * ctx->n];
                }
                f.q = emit->stats_q; f.ldq = emit->stats_ldq; f.slot0 = emit->slot0; f.nch = emit->nch;
                torus_stats_finish_kernel<<<(unsigned)((emit->nch + 127) / 128), 128, 0, ctx->stream>>>(f);
                MCL_CUDA(ctx, cudaGetLastError());
            }
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

int torus_emit(mclcar_ctx* ctx, TorusChains& t, const double* d_mu, double mu_const, void* out, int dtype, int64_t ld,
               int64_t slot0, int64_t nch) {
    const int h = ctx->n2 / 2;
    dim3 block(128);
    dim3 grid((h + 127) / 128, ctx->n1, (unsigned)nch);
    ProfScope prof(ctx, PROF_EMIT, 1);
    if (dtype == MCLCAR_F32)
        torus_emit_kernel<float><<<grid, block, 0, ctx->stream>>>(t.red, t.black, ctx->n1, h, d_mu, mu_const, (float*)out, ld, slot0, nch);
    else
        torus_emit_kernel<double><<<grid, block, 0, ctx->stream>>>(t.red, t.black, ctx->n1, h, d_mu, mu_const, (double*)out, ld, slot0, nch);
    MCL_CUDA(ctx, cudaGetLastError());
    return MCLCAR_OK;
}

}  // namespace mclcar

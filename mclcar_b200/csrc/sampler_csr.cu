// K1 (general graph) -- batched chromatic Gibbs / Metropolis-within-Gibbs sampler for a
// Gaussian Markov random field  pi(x) ~ exp(-x'Qx/2 + b'x) [ * prod_s L_s(x_s) ],  Q sparse.
//
//   GAUSS : x_s | rest ~ N(m_s, 1/Q_ss),  m_s = (b_s - sum_{j != s} Q_sj x_j)/Q_ss        (exact Gibbs)
//           - direct CAR on a neighbourhood matrix W: Q = (I - rho W)/sigma, b = 0, sample = X beta0 + x.
//             Replaces CAR.simWmat's Cholesky draw (R/CAR.simLM.R:25-34).
//           - hierarchical CAR: Q = Z'Qe Z + Qu, b = Z'Qe (y - X beta): the u | y draw of sim.HCAR
//             (R/HCAR.sim.R:64-69).
//   BINOM / POISSON : latent CAR field of the GLM, target log pi(z) = y'z - sum n log(1+e^{xb+z})
//           - z'Qz/2 (R/postZsub.R:431) | y'z - sum e^{xb+z} - z'Qz/2 (:976).  Per site a Metropolis-Hastings
//           step: two sweeps in three propose from a Gaussian approximation of the site's full conditional
//           (independence proposal, glm_laplace below), the third from the prior conditional N(m_s, 1/Q_ss),
//           for which the ratio is the likelihood ratio only.  Replaces the single-chain whole-vector MALA of
//           postZ (R/postZsub.R:393-529).
//
// Parallelisation: chains are independent, so one CTA owns CB chains for the WHOLE run
// (burn-in, thinning, all kept states) with their state resident in shared memory, [site][CB]
// floats; colour classes are separated by __syncthreads() only -- no grid-wide barrier and
// one launch per sample set.  One thread = (site slot, chain pair): the pair's two normals
// and two uniforms come from ONE Philox4x32-10 call keyed by (seed; site, sweep, global pair
// id).  Kept states are written site-major ("row = sample" of the R matrix Zy): for every
// site the CB chains of a CTA form one contiguous run.  Maps whose neighbour lists fit behind the
// state keep the graph on chip too (MODE 2: 16-bit neighbour ids, 8-byte position records).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "engine.hpp"

namespace mclcar {

struct alignas(16) CsrSite {
    int s;         // site id
    int e0, e1;    // its (padded) neighbour entries
    float hm;      // b_s / Q_ss
    float sd;      // 1 / sqrt(Q_ss)
    float y, nt, xb;   // GLM data of the site (observation, trials, X beta)
};

struct CsrSamplerParams {
    int n, ncolors;
    const int* color_ptr;  // [ncolors + 1] positions
    const struct CsrSite* meta;  // [n] by position: everything a site update reads besides its neighbours, in one 32-byte record
    const int* colidx;     // neighbour SITE ids; every row padded to a multiple of four entries with the dummy site n,
                           // whose state row stays zero: the gather loop runs on whole int4 / float4 groups, no predicates
    const float* coef;     // -Q_sj / Q_ss per entry (0 for the padding)
    int coef_uniform;      // all entries equal (binary W, constant diagonal): use coef_value, skip the loads
    float coef_value;
    const double* mu;      // added to kept states (site order) or null
    int CB;            // chains per CTA (even)
    int64_t C;         // chains on this context = CB * gridDim.x
    int burn, thin, E; // E kept states per chain
    int64_t N;         // samples wanted on this context (<= E * C)
    void* out;         // [n][ld] site-major, float or double
    int64_t ld;
    int out_f64;
    float* state_g;    // global-memory state [grid][n][CB] when it does not fit in smem, else null
    PhiloxKeys keys;
    float keep, gain, noise;   // over-relaxation: x' = keep*x + gain*m + noise*sd*z (Gaussian family)
    int64_t pair0, pair_stride;  // global pair id = pair0 + local_pair * pair_stride
    unsigned long long* counters;  // [0] accepted, [1] proposed
    // statistics of the kept states formed at emission, where the state and all its neighbours are on chip:
    // stat_q[i] = x_i'x_i, stat_q[stat_ldq + i] = x_i'W x_i for a BINARY neighbourhood matrix (every real neighbour entry
    // counts once; the padding's dummy site holds zeros), in fp64 from the fp32 values that are written out -- what the
    // site-major statistics kernel would compute from the stored matrix, without its second pass over it.  null: off
    double* stat_q;
    int64_t stat_ldq;
    // graph on chip (MODE 2): the padded neighbour lists as 16-bit site ids, four to an 8-byte group, and one 8-byte record
    // per position (site | groups << 16, first group) are copied into shared memory behind the state, so that a site
    // update makes ONE global load (the second half of its CsrSite record) that nothing else waits for
    const uint2* idx16;    // [ngroups]
    const uint2* site8;    // [n]
    int ngroups;
    int icm;               // start-up sweeps that move every site to the mode of its approximation (never more than the burn-in)
    int laplace;           // GLM families: 1 = proposals from the Gaussian approximation of the full conditional, 0 = from the prior conditional
    uint32_t scratch_off;  // byte offset of the statistics scratch in shared memory
};

namespace {

constexpr uint32_t kTagCsr = 0x4373u;
constexpr int kMaxThreads = 1024;

__device__ __forceinline__ float softplusf(float t) { return fmaxf(t, 0.f) + __logf(1.f + __expf(-fabsf(t))); }

// log-likelihood difference of one site, f(a) - f(b):  f(z) = y z - n log(1 + e^{xb + z})  |  y z - e^{xb + z}
template <int FAM>
__device__ __forceinline__ float glm_logf(const float a, const float b, const float ys, const float ns, const float xbs) {
    if (FAM == FAM_BINOM) return ys * (a - b) - ns * (softplusf(xbs + a) - softplusf(xbs + b));
    return ys * (a - b) - (__expf(xbs + a) - __expf(xbs + b));
}

// Gaussian approximation N(mu, kInflate / tau) of the full conditional of one latent site,
//   log pi(z) = f(z) - tau0 (z - m)^2 / 2,   m, 1 / tau0: mean and variance of its prior conditional:
// mu = the mode, tau = tau0 - f''(mu), found by two Newton steps from the precision-weighted average of m and the
// likelihood's own mode zb (curvature d2b there; both depend on the site's data only).  Steps are limited to four
// standard deviations of the current approximation.  The approximation depends on the neighbours and the data, not on
// the site's current value: proposing from it is an independence Metropolis-Hastings step with acceptance
// min(1, w(z') / w(z)), w = pi / q -- close to one where the likelihood is sharp, which is where proposals from the
// prior conditional are rarely accepted (config 3: 34 %).  An independence sampler is only as good as the tails of its
// proposal: the variance is inflated a little, and every third sweep proposes from the prior conditional instead, whose
// tails are those of the target (sites with extreme counts have a one-sided flat likelihood).
constexpr float kInflate = 1.25f;
constexpr int kIcmSweeps = 4;
template <int FAM>
__device__ __forceinline__ void glm_laplace(const float m, const float tau0, const float ys, const float ns, const float xbs,
                                            const float zb, const float d2b, float& mu, float& tau) {
    float z0 = __fdividef(fmaf(tau0, m, d2b * zb), tau0 + d2b);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        float d1, d2;   // f'(z0), -f''(z0)
        if (FAM == FAM_BINOM) {
            const float pz = __frcp_rn(1.f + __expf(-(xbs + z0)));
            d1 = fmaf(-ns, pz, ys);
            d2 = ns * pz * (1.f - pz);
        } else {
            const float ee = __expf(fminf(xbs + z0, 80.f));
            d1 = ys - ee;
            d2 = ee;
        }
        tau = tau0 + d2;
        const float lim = 4.f * rsqrtf(tau);
        z0 += fminf(fmaxf(__fdividef(fmaf(tau0, m - z0, d1), tau), -lim), lim);
    }
    mu = z0;
    tau *= 1.f / kInflate;
}

// The chain state of a CTA, [site][CB] floats: in shared memory (SMEM: explicit 32-bit shared-space addresses, so the
// gathers are LDS.64 with one integer multiply-add each -- through a pointer that may be shared OR global every access
// was a generic load behind 64-bit address arithmetic) or, for graphs too large for that, in global memory.
template <bool SMEM>
struct ChainState {
    uint32_t sbase;      // shared-space byte address of the state (SMEM)
    float* g;            // global state of this CTA (!SMEM)
    __device__ __forceinline__ float2 ld2(const int elem) const {   // elem = site * CB + 2 * pair
        if (SMEM) {
            float2 v;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sbase + 4u * (uint32_t)elem));
            return v;
        }
        return *reinterpret_cast<const float2*>(g + elem);
    }
    __device__ __forceinline__ void st2(const int elem, const float2 v) const {
        if (SMEM) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sbase + 4u * (uint32_t)elem), "f"(v.x), "f"(v.y) : "memory");
        else *reinterpret_cast<float2*>(g + elem) = v;
    }
    __device__ __forceinline__ void zero(const int elem) const {
        if (SMEM) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbase + 4u * (uint32_t)elem), "f"(0.f) : "memory");
        else g[elem] = 0.f;
    }
};

__device__ __forceinline__ uint2 lds_u2(const uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u2(const uint32_t a, const uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}

// MODE: 0 chain state in global memory, 1 in shared memory, 2 in shared memory together with the graph
template <int FAM, int MODE>
__global__ void __launch_bounds__(kMaxThreads) gibbs_csr_kernel(const CsrSamplerParams p) {
    constexpr bool SMEM = MODE >= 1, GOC = MODE == 2;
    const int kThreads = blockDim.x;
    extern __shared__ float sm_state[];
    const int CB = p.CB, hp = CB / 2;           // pairs per site row
    const int pr = threadIdx.x % hp;            // chain pair within the CTA
    const int slot = threadIdx.x / hp;
    const int nslots = kThreads / hp;
    ChainState<SMEM> x;
    x.sbase = smem_u32(sm_state);
    x.g = SMEM ? nullptr : p.state_g + (size_t)blockIdx.x * (p.n + 1) * CB;   // row n: the dummy site of the padding, always zero
    for (int t = threadIdx.x; t < (p.n + 1) * CB; t += kThreads) x.zero(t);
    const uint32_t idx_base = x.sbase + (((uint32_t)(p.n + 1) * (uint32_t)CB * 4u + 7u) & ~7u);
    const uint32_t site_base = idx_base + 8u * (uint32_t)p.ngroups;
    if (GOC) {
        for (int t = threadIdx.x; t < p.ngroups; t += kThreads) sts_u2(idx_base + 8u * (uint32_t)t, __ldg(p.idx16 + t));
        for (int t = threadIdx.x; t < p.n; t += kThreads) sts_u2(site_base + 8u * (uint32_t)t, __ldg(p.site8 + t));
    }
    __syncthreads();

    const uint64_t gpair = (uint64_t)(p.pair0 + ((int64_t)blockIdx.x * hp + pr) * p.pair_stride);
    const int64_t chain_local = (int64_t)blockIdx.x * CB + 2 * pr;
    unsigned int acc_cnt = 0, prop_cnt = 0;
    const int total = p.burn + p.E * p.thin;
    int e_next = 0;
    for (int sweep = 0; sweep < total; ++sweep) {
        for (int col = 0; col < p.ncolors; ++col) {
            const int t1 = p.color_ptr[col + 1];
            for (int t = p.color_ptr[col] + slot; t < t1; t += nslots) {
                const float4 mb = __ldg(reinterpret_cast<const float4*>(p.meta + t) + 1);  // sd, y (Gaussian family: b_s / Q_ss), nt, xb
                int s;
                float m0, m1;
                if constexpr (GOC) {
                    const uint2 rec = lds_u2(site_base + 8u * (uint32_t)t);
                    s = (int)(rec.x & 0xffffu);
                    m0 = m1 = FAM == FAM_GAUSS ? mb.y : 0.f;
                    const uint32_t pbase = x.sbase + 8u * (uint32_t)pr, cb4 = 4u * (uint32_t)CB;
                    const uint32_t gend = rec.y + (rec.x >> 16);
                    for (uint32_t g = rec.y; g < gend; ++g) {
                        const uint2 jj = lds_u2(idx_base + 8u * g);
                        float4 c;
                        if (p.coef_uniform) c = make_float4(p.coef_value, p.coef_value, p.coef_value, p.coef_value);
                        else c = __ldg(reinterpret_cast<const float4*>(p.coef) + g);
                        float2 v0, v1, v2, v3;
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v0.x), "=f"(v0.y) : "r"((jj.x & 0xffffu) * cb4 + pbase));
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v1.x), "=f"(v1.y) : "r"((jj.x >> 16) * cb4 + pbase));
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v2.x), "=f"(v2.y) : "r"((jj.y & 0xffffu) * cb4 + pbase));
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v3.x), "=f"(v3.y) : "r"((jj.y >> 16) * cb4 + pbase));
                        m0 = fmaf(c.x, v0.x, m0); m1 = fmaf(c.x, v0.y, m1);
                        m0 = fmaf(c.y, v1.x, m0); m1 = fmaf(c.y, v1.y, m1);
                        m0 = fmaf(c.z, v2.x, m0); m1 = fmaf(c.z, v2.y, m1);
                        m0 = fmaf(c.w, v3.x, m0); m1 = fmaf(c.w, v3.y, m1);
                    }
                } else {
                    const int4 ma = __ldg(reinterpret_cast<const int4*>(p.meta + t));          // s, e0, e1, hm
                    s = ma.x;
                    m0 = m1 = __int_as_float(ma.w);
                    // neighbour gathers four at a time (rows are padded to whole groups): one 16-byte index load, four
                    // independent index -> shared-memory chains in flight
                    for (int e = ma.y; e < ma.z; e += 4) {
                        const int4 j = __ldg(reinterpret_cast<const int4*>(p.colidx + e));
                        float4 c;
                        if (p.coef_uniform) c = make_float4(p.coef_value, p.coef_value, p.coef_value, p.coef_value);
                        else c = __ldg(reinterpret_cast<const float4*>(p.coef + e));
                        const float2 v0 = x.ld2(j.x * CB + 2 * pr), v1 = x.ld2(j.y * CB + 2 * pr);
                        const float2 v2 = x.ld2(j.z * CB + 2 * pr), v3 = x.ld2(j.w * CB + 2 * pr);
                        m0 = fmaf(c.x, v0.x, m0); m1 = fmaf(c.x, v0.y, m1);
                        m0 = fmaf(c.y, v1.x, m0); m1 = fmaf(c.y, v1.y, m1);
                        m0 = fmaf(c.z, v2.x, m0); m1 = fmaf(c.z, v2.y, m1);
                        m0 = fmaf(c.w, v3.x, m0); m1 = fmaf(c.w, v3.y, m1);
                    }
                }
                const Philox4 r = philox4x32_10((uint32_t)s, (uint32_t)sweep, (uint32_t)gpair,
                                                (uint32_t)(gpair >> 32) ^ (kTagCsr << 16), p.keys);
                float z0, z1;
                box_muller(r.x, r.y, z0, z1);
                const float sdv = mb.x;
                float2 nv;
                if constexpr (FAM == FAM_GAUSS) {
                    const float2 cur = x.ld2(s * CB + 2 * pr);
                    nv.x = fmaf(p.keep, cur.x, fmaf(p.gain, m0, p.noise * sdv * z0));
                    nv.y = fmaf(p.keep, cur.y, fmaf(p.gain, m1, p.noise * sdv * z1));
                }
                if constexpr (FAM != FAM_GAUSS) {
                    const float2 cur = x.ld2(s * CB + 2 * pr);
                    const float ys = mb.y, ns = mb.z, xbs = mb.w;
                    if (p.laplace > 0 && (sweep < p.icm || p.laplace == 1 || sweep % p.laplace != p.laplace - 1)) {
                        // independence proposal from a Gaussian approximation of the site's full conditional (see glm_laplace)
                        const float tau0 = __frcp_rn(sdv * sdv);
                        float zb, d2b;          // mode and curvature of the site's likelihood alone
                        if (FAM == FAM_BINOM) {
                            const float pb = __fdividef(ys + 0.5f, ns + 1.f);
                            zb = __logf(__fdividef(pb, 1.f - pb)) - xbs;
                            d2b = ns * pb * (1.f - pb);
                        } else {
                            zb = __logf(ys + 0.5f) - xbs;
                            d2b = ys + 0.5f;
                        }
                        float mu0, mu1, t0, t1;
                        glm_laplace<FAM>(m0, tau0, ys, ns, xbs, zb, d2b, mu0, t0);
                        glm_laplace<FAM>(m1, tau0, ys, ns, xbs, zb, d2b, mu1, t1);
                        nv.x = fmaf(rsqrtf(t0), z0, mu0);
                        nv.y = fmaf(rsqrtf(t1), z1, mu1);
                        if (sweep < p.icm) {
                            // start-up: the chains begin at zero, which for a site with a large count lies tens of
                            // standard deviations out in the tail of its approximation -- a state no independence
                            // proposal is ever accepted from.  The first sweeps move every site to the mode of its
                            // conditional instead (iterated conditional modes); the burn-in proper starts from there.
                            x.st2(s * CB + 2 * pr, make_float2(mu0, mu1));
                            continue;
                        }
                        const float la0 = glm_logf<FAM>(nv.x, cur.x, ys, ns, xbs) +
                                          0.5f * (tau0 * ((cur.x - m0) * (cur.x - m0) - (nv.x - m0) * (nv.x - m0)) + z0 * z0 - t0 * (cur.x - mu0) * (cur.x - mu0));
                        const float la1 = glm_logf<FAM>(nv.y, cur.y, ys, ns, xbs) +
                                          0.5f * (tau0 * ((cur.y - m1) * (cur.y - m1) - (nv.y - m1) * (nv.y - m1)) + z1 * z1 - t1 * (cur.y - mu1) * (cur.y - mu1));
                        const bool a0 = __logf(u01(r.z)) < la0, a1 = __logf(u01(r.w)) < la1;
                        if (!a0) nv.x = cur.x;
                        if (!a1) nv.y = cur.y;
                        acc_cnt += (a0 ? 1u : 0u) + (a1 ? 1u : 0u);
                    } else {
                        // proposal = the prior conditional N(m, 1 / Q_ss): the Metropolis ratio is the likelihood ratio
                        nv.x = fmaf(sdv, z0, m0);
                        nv.y = fmaf(sdv, z1, m1);
                        const float la0 = glm_logf<FAM>(nv.x, cur.x, ys, ns, xbs), la1 = glm_logf<FAM>(nv.y, cur.y, ys, ns, xbs);
                        const bool a0 = __logf(u01(r.z)) < la0, a1 = __logf(u01(r.w)) < la1;
                        if (!a0) nv.x = cur.x;
                        if (!a1) nv.y = cur.y;
                        acc_cnt += (a0 ? 1u : 0u) + (a1 ? 1u : 0u);
                    }
                    prop_cnt += 2u;
                }
                x.st2(s * CB + 2 * pr, nv);
            }
            __syncthreads();
        }
        if (sweep + 1 == p.burn + (e_next + 1) * p.thin) {
            // keep this state: sample index = e * C + chain
            const int64_t i0 = (int64_t)e_next * p.C + chain_local;
            for (int s = slot; s < p.n; s += nslots) {
                const float2 v = x.ld2(s * CB + 2 * pr);
                const double m = p.mu ? p.mu[s] : 0.0;
                if (p.out_f64) {
                    double* o = reinterpret_cast<double*>(p.out) + (size_t)s * p.ld;
                    if (i0 < p.N) o[i0] = (double)v.x + m;
                    if (i0 + 1 < p.N) o[i0 + 1] = (double)v.y + m;
                } else {
                    float* o = reinterpret_cast<float*>(p.out) + (size_t)s * p.ld;
                    if (i0 < p.N) o[i0] = (float)((double)v.x + m);
                    if (i0 + 1 < p.N) o[i0 + 1] = (float)((double)v.y + m);
                }
            }
            if (p.stat_q != nullptr) {
                // x'x and x'Wx of the state just written: every thread takes its share of the sites for its chain pair,
                // lanes of a warp holding the same pair are added by shuffles, the warps through shared memory
                double zz0 = 0.0, zz1 = 0.0, zw0 = 0.0, zw1 = 0.0;
                for (int t = slot; t < p.n; t += nslots) {
                    const int4 ma = __ldg(reinterpret_cast<const int4*>(p.meta + t));
                    const float2 v = x.ld2(ma.x * CB + 2 * pr);
                    double g0 = 0.0, g1 = 0.0;
                    for (int e = ma.y; e < ma.z; e += 4) {
                        const int4 j = __ldg(reinterpret_cast<const int4*>(p.colidx + e));
                        const float2 v0 = x.ld2(j.x * CB + 2 * pr), v1 = x.ld2(j.y * CB + 2 * pr);
                        const float2 v2 = x.ld2(j.z * CB + 2 * pr), v3 = x.ld2(j.w * CB + 2 * pr);
                        g0 += ((double)v0.x + (double)v1.x) + ((double)v2.x + (double)v3.x);
                        g1 += ((double)v0.y + (double)v1.y) + ((double)v2.y + (double)v3.y);
                    }
                    zz0 = fma((double)v.x, (double)v.x, zz0);
                    zz1 = fma((double)v.y, (double)v.y, zz1);
                    zw0 = fma((double)v.x, g0, zw0);
                    zw1 = fma((double)v.y, g1, zw1);
                }
                for (int o = hp; o < 32; o <<= 1) {          // lanes l, l + hp, l + 2 hp, ... hold the same pair (hp <= 16)
                    zz0 += __shfl_xor_sync(0xffffffffu, zz0, o); zz1 += __shfl_xor_sync(0xffffffffu, zz1, o);
                    zw0 += __shfl_xor_sync(0xffffffffu, zw0, o); zw1 += __shfl_xor_sync(0xffffffffu, zw1, o);
                }
                double* sc = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sm_state) + p.scratch_off);   // [warps][hp][4]
                const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = kThreads >> 5;
                __syncthreads();
                if (lane < hp) {
                    double* d = sc + ((size_t)wid * hp + lane) * 4;
                    d[0] = zz0; d[1] = zz1; d[2] = zw0; d[3] = zw1;
                }
                __syncthreads();
                if (threadIdx.x < hp) {                       // thread = pair: add the warps in order (deterministic)
                    double a[4] = {0.0, 0.0, 0.0, 0.0};
                    for (int w = 0; w < nw; ++w) {
                        const double* d = sc + ((size_t)w * hp + threadIdx.x) * 4;
                        a[0] += d[0]; a[1] += d[1]; a[2] += d[2]; a[3] += d[3];
                    }
                    const int64_t i0s = (int64_t)e_next * p.C + (int64_t)blockIdx.x * CB + 2 * threadIdx.x;
                    if (i0s < p.N) { p.stat_q[i0s] = a[0]; p.stat_q[p.stat_ldq + i0s] = a[2]; }
                    if (i0s + 1 < p.N) { p.stat_q[i0s + 1] = a[1]; p.stat_q[p.stat_ldq + i0s + 1] = a[3]; }
                }
                __syncthreads();
            }
            ++e_next;
        }
    }
    if constexpr (FAM != FAM_GAUSS) {
        // block totals -> two atomics per CTA
        __shared__ unsigned int s_acc, s_prop;
        if (threadIdx.x == 0) s_acc = s_prop = 0;
        __syncthreads();
        atomicAdd(&s_acc, acc_cnt);
        atomicAdd(&s_prop, prop_cnt);
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(p.counters, (unsigned long long)s_acc);
            atomicAdd(p.counters + 1, (unsigned long long)s_prop);
        }
    }
}

}  // namespace

// A colour class of size m occupies ceil(m / nslots) rounds of the CTA's site slots, so a sweep costs
// sum_c ceil(size_c / nslots) rounds whatever the total: first-fit colouring leaves a long tail of small classes and
// large ones just over a multiple of nslots (config-3 map, 512 slots: 1519 + 1419 + 1256 + 740 + 66 sites = 12 rounds
// against ceil(5000 / 512) = 10).  Move the overflow of a class beyond its last full round into the spare room of
// the other classes where the neighbourhoods allow it (any proper colouring gives a valid chromatic sweep).
// order / cptr: colour-major site order and class boundaries, rewritten when a colouring with fewer rounds was found.
static bool rebalance_colours(const GmrfHost& g, int nslots, std::vector<int>& order, std::vector<int>& cptr) {
    const int n = g.n;
    int ncol = (int)cptr.size() - 1;
    if (nslots < 1 || ncol < 2) return false;
    std::vector<int> col(n), size(ncol), m(ncol);
    for (int c = 0; c < ncol; ++c) {
        size[c] = cptr[c + 1] - cptr[c];
        m[c] = (size[c] + nslots - 1) / nslots;
        for (int t = cptr[c]; t < cptr[c + 1]; ++t) col[order[t]] = c;
    }
    auto rounds = [&]() { int r = 0; for (int c = 0; c < ncol; ++c) r += m[c]; return r; };
    const int before = rounds(), ideal = (n + nslots - 1) / nslots;
    if (before <= ideal) return false;
    std::vector<char> seen(ncol);
    for (int pass = 0; pass < 4 && rounds() > ideal; ++pass) {
        bool improved = false;
        std::vector<int> byover(ncol);
        for (int c = 0; c < ncol; ++c) byover[c] = c;
        auto over = [&](int c) { return size[c] - (m[c] - 1) * nslots; };   // sites in the last round of class c
        std::sort(byover.begin(), byover.end(), [&](int a, int b) { return over(a) < over(b); });
        for (int c : byover) {
            if (m[c] < 1 || rounds() <= ideal) continue;
            int need = over(c);
            int room = 0;
            for (int d = 0; d < ncol; ++d) if (d != c) room += m[d] * nslots - size[d];
            if (need > room) continue;
            std::vector<std::pair<int, int>> moved;   // (site, new class)
            for (int i = 0; i < n && need > 0; ++i) {
                if (col[i] != c) continue;
                std::fill(seen.begin(), seen.end(), 0);
                for (int e = (*g.rowptr)[i]; e < (*g.rowptr)[i + 1]; ++e) seen[col[(*g.colidx)[e]]] = 1;
                int best = -1, best_room = 0;
                for (int d = 0; d < ncol; ++d) {
                    const int r = m[d] * nslots - size[d];
                    if (d != c && !seen[d] && r > best_room) { best = d; best_room = r; }
                }
                if (best < 0) continue;
                col[i] = best; ++size[best]; --size[c]; --need;
                moved.push_back({i, best});
            }
            if (need > 0) {   // not enough movable sites: put them back
                for (auto& mv : moved) { col[mv.first] = c; --size[mv.second]; ++size[c]; }
            } else {
                --m[c];
                improved = true;
            }
        }
        if (!improved) break;
    }
    if (rounds() >= before) return false;
    std::vector<int> new_order, new_ptr(1, 0);
    new_order.reserve(n);
    for (int c = 0; c < ncol; ++c) {
        if (!size[c]) continue;
        for (int i = 0; i < n; ++i) if (col[i] == c) new_order.push_back(i);
        new_ptr.push_back((int)new_order.size());
    }
    order.swap(new_order);
    cptr.swap(new_ptr);
    return true;
}

int run_csr_sampler(mclcar_ctx* ctx, const GmrfHost& g, int fam, double omega, const double* mu_host, const double* y,
                    const double* ntrial, const double* xb, int64_t N, int64_t C_want, int burn, int thin,
                    void* d_out, int out_f64, int64_t ld, double* accept_rate, int64_t* sweeps_out, double* d_stat_q, int64_t stat_ldq,
                    bool* stats_done) {
    MCL_TRY(require_device(ctx));
    if (stats_done) *stats_done = false;
    begin_draw(ctx);   // a fresh Philox stream for every sampler run (the reference draws fresh rnorm() per call)
    const int n = g.n;
    // ---- chains per CTA: largest CB whose state fits in shared memory and still fills the SMs
    const size_t smem_cap = 200 * 1024;
    int CB = 0;
    for (int cb : {32, 16, 8, 4, 2})
        if ((size_t)(n + 1) * cb * sizeof(float) <= smem_cap) {
            CB = cb;
            if ((C_want + cb - 1) / cb >= ctx->n_sm) break;
        }
    const bool in_smem = CB != 0;
    if (!in_smem) CB = 32;
    int64_t grid = (C_want + CB - 1) / CB;
    if (grid < 1) grid = 1;
    const int64_t C = grid * CB;
    int E = (int)((N + C - 1) / C);
    if (E < 1) E = 1;
    // all warps of the CTA share the resident state: as many as the largest colour class can feed
    std::vector<int> order(*g.order), cptr(*g.color_ptr);
    int max_class = 1;
    for (size_t c = 0; c + 1 < cptr.size(); ++c) max_class = std::max(max_class, cptr[c + 1] - cptr[c]);
    int threads = 256;
    while (threads < kMaxThreads && (threads / (CB / 2)) < max_class) threads *= 2;
    if (const char* te = getenv("MCLCAR_CSR_THREADS")) threads = atoi(te);
    {   // colour classes sized for whole rounds of the CTA's site slots (MCLCAR_CSR_REBALANCE=0: the graph's own colouring)
        const char* re = getenv("MCLCAR_CSR_REBALANCE");
        if (!(re && atoi(re) == 0)) rebalance_colours(g, threads / (CB / 2), order, cptr);
    }
    const int ncol = (int)cptr.size() - 1;
    // ---- permute the CSR into colour-major position order, convert to the conditional form
    // (rows padded to whole groups of four entries with the dummy site n, coefficient 0) and one 32-byte record per
    // position with everything else a site update reads
    std::vector<int> ci;
    std::vector<float> coef;
    std::vector<CsrSite> meta(n);
    ci.reserve(g.colidx->size() + 3 * (size_t)n);
    coef.reserve(g.colidx->size() + 3 * (size_t)n);
    bool coef_uniform = true;
    float coef_first = 0.f;
    bool have_first = false;
    for (int t = 0; t < n; ++t) {
        const int s = order[t];
        const double qss = g.qdiag[s];
        if (!(qss > 0)) return fail(ctx, MCLCAR_EINVAL, "precision matrix has a non-positive diagonal at site %d", s);
        CsrSite& ms = meta[t];
        ms.s = s;
        ms.e0 = (int)ci.size();
        for (int e = (*g.rowptr)[s]; e < (*g.rowptr)[s + 1]; ++e) {
            const float c = (float)(-g.qoff[e] / qss);
            ci.push_back((*g.colidx)[e]);
            coef.push_back(c);
            if (!have_first) { coef_first = c; have_first = true; }
            coef_uniform = coef_uniform && c == coef_first;
        }
        while (ci.size() % 4) { ci.push_back(n); coef.push_back(0.f); }
        ms.e1 = (int)ci.size();
        ms.hm = g.b.empty() ? 0.f : (float)(g.b[s] / qss);
        ms.sd = (float)(1.0 / std::sqrt(qss));
        ms.y = fam != FAM_GAUSS ? (float)y[s] : ms.hm;   // Gaussian family: the second half of the record carries b_s / Q_ss too
        ms.nt = fam != FAM_GAUSS ? (ntrial ? (float)ntrial[s] : 1.f) : 0.f;
        ms.xb = fam != FAM_GAUSS ? (float)xb[s] : 0.f;
    }
    // ---- upload
    auto up = [&](const void* h, size_t bytes, void** d) -> int {
        MCL_TRY(pool_alloc(ctx, d, bytes ? bytes : 4));
        if (bytes) MCL_CUDA(ctx, cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return MCLCAR_OK;
    };
    std::vector<void*> owned;
    auto upv = [&](const void* h, size_t bytes, void** d) -> int {
        int st = up(h, bytes, d);
        if (st == MCLCAR_OK) owned.push_back(*d);
        return st;
    };
    CsrSamplerParams p{};
    int st = MCLCAR_OK;
    std::vector<double> mu;
    void *d_cp = nullptr, *d_meta = nullptr, *d_ci = nullptr, *d_coef = nullptr, *d_mu = nullptr, *d_cnt = nullptr, *d_state = nullptr;
    do {
        if ((st = upv(cptr.data(), (ncol + 1) * sizeof(int), &d_cp))) break;
        if ((st = upv(meta.data(), meta.size() * sizeof(CsrSite), &d_meta))) break;
        if ((st = upv(ci.data(), ci.size() * sizeof(int), &d_ci))) break;
        if ((st = upv(coef.data(), coef.size() * sizeof(float), &d_coef))) break;
        if (mu_host) {
            if ((st = upv(mu_host, n * sizeof(double), &d_mu))) break;
        }
        unsigned long long zero[2] = {0, 0};
        if ((st = upv(zero, sizeof zero, &d_cnt))) break;
        if (!in_smem) {
            if ((st = pool_alloc(ctx, &d_state, (size_t)grid * (n + 1) * CB * sizeof(float)))) break;
            owned.push_back(d_state);
        }
        p.n = n; p.ncolors = ncol;
        p.color_ptr = (const int*)d_cp; p.meta = (const CsrSite*)d_meta;
        p.colidx = (const int*)d_ci; p.coef = (const float*)d_coef;
        p.coef_uniform = coef_uniform ? 1 : 0;
        p.coef_value = coef_first;
        p.mu = (const double*)d_mu;
        p.CB = CB; p.C = C; p.burn = burn; p.thin = thin; p.E = E; p.N = N;
        p.out = d_out; p.ld = ld; p.out_f64 = out_f64; p.state_g = (float*)d_state;
        p.keys = philox_keys(ctx->draw_key);
        p.keep = (float)(1.0 - omega); p.gain = (float)omega; p.noise = (float)std::sqrt(omega * (2.0 - omega));
        p.pair0 = ctx->rank; p.pair_stride = ctx->world;
        p.counters = (unsigned long long*)d_cnt;
        {   // 0: proposals from the prior conditional; 1: from the Gaussian approximation; k > 1: every k-th sweep from the prior
            const char* le = getenv("MCLCAR_GLM_PROPOSAL");
            p.laplace = 3;
            if (le && !strcmp(le, "prior")) p.laplace = 0;
            else if (le && !strncmp(le, "laplace", 7) && le[7]) p.laplace = std::max(1, atoi(le + 7));
            p.icm = p.laplace > 0 ? std::min(kIcmSweeps, burn) : 0;
        }
        // statistics at emission: binary neighbourhood matrices, no mean added to the kept states, fp32 output (the
        // statistics must be those of the values written)
        const bool on_chip_stats = d_stat_q != nullptr && coef_uniform && have_first && !mu_host && !out_f64;
        if (on_chip_stats) { p.stat_q = d_stat_q; p.stat_ldq = stat_ldq; }
        if (stats_done) *stats_done = on_chip_stats;
        size_t smem = in_smem ? (((size_t)(n + 1) * CB * sizeof(float) + 7) & ~(size_t)7) : 0;
        const size_t stat_bytes = on_chip_stats ? (size_t)(threads / 32) * (CB / 2) * 4 * sizeof(double) : 0;
        // graph on chip: 16-bit neighbour ids and the 8-byte position records behind the state, when everything fits
        const size_t ngroups = ci.size() / 4;
        const size_t graph_bytes = ngroups * 8 + (size_t)n * 8;
        const char* ge = getenv("MCLCAR_CSR_GRAPH_ON_CHIP");
        const bool goc = in_smem && n <= 65535 && ((size_t)n * sizeof(CsrSite) + ci.size() * 4 > 48 * 1024 || (ge && atoi(ge) == 1)) && ngroups <= 0x7fffffffu && (fam == FAM_GAUSS || g.b.empty()) &&
                         smem + graph_bytes + stat_bytes <= 220 * 1024 && !(ge && atoi(ge) == 0);
        if (goc) {
            std::vector<uint2> idx16(ngroups), site8(n);
            bool ok = true;
            for (size_t q = 0; q < ngroups; ++q) {
                idx16[q].x = (uint32_t)ci[4 * q] | ((uint32_t)ci[4 * q + 1] << 16);
                idx16[q].y = (uint32_t)ci[4 * q + 2] | ((uint32_t)ci[4 * q + 3] << 16);
            }
            for (int t = 0; t < n; ++t) {
                const uint32_t ng = (uint32_t)(meta[t].e1 - meta[t].e0) / 4;
                ok = ok && ng <= 0xffffu;
                site8[t].x = (uint32_t)meta[t].s | (ng << 16);
                site8[t].y = (uint32_t)meta[t].e0 / 4;
            }
            if (!ok) { st = fail(ctx, MCLCAR_ELIMIT, "a site has more than 262140 neighbours"); break; }
            void *d_i16 = nullptr, *d_s8 = nullptr;
            if ((st = upv(idx16.data(), idx16.size() * sizeof(uint2), &d_i16))) break;
            if ((st = upv(site8.data(), site8.size() * sizeof(uint2), &d_s8))) break;
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { st = fail(ctx, MCLCAR_ECUDA, "upload of the on-chip graph failed"); break; }   // the two vectors are locals
            p.idx16 = (const uint2*)d_i16; p.site8 = (const uint2*)d_s8; p.ngroups = (int)ngroups;
            smem += graph_bytes;
        }
        p.scratch_off = (uint32_t)smem;
        smem += stat_bytes;   // scratch of the statistics
        cudaError_t e = cudaSuccess;
        ProfScope prof(ctx, PROF_CSR_SAMPLER, 1);
#define LAUNCH_CSR(F_, S_)                                                                                              \
    do {                                                                                                                \
        e = cudaFuncSetAttribute(gibbs_csr_kernel<F_, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_cap + 24 * 1024));  \
        if (e == cudaSuccess) gibbs_csr_kernel<F_, S_><<<(unsigned)grid, threads, smem, ctx->stream>>>(p);              \
    } while (0)
        const int mode = goc ? 2 : (in_smem ? 1 : 0);
        if (fam == FAM_GAUSS) { if (mode == 2) LAUNCH_CSR(FAM_GAUSS, 2); else if (mode == 1) LAUNCH_CSR(FAM_GAUSS, 1); else LAUNCH_CSR(FAM_GAUSS, 0); }
        else if (fam == FAM_BINOM) { if (mode == 2) LAUNCH_CSR(FAM_BINOM, 2); else if (mode == 1) LAUNCH_CSR(FAM_BINOM, 1); else LAUNCH_CSR(FAM_BINOM, 0); }
        else { if (mode == 2) LAUNCH_CSR(FAM_POIS, 2); else if (mode == 1) LAUNCH_CSR(FAM_POIS, 1); else LAUNCH_CSR(FAM_POIS, 0); }
#undef LAUNCH_CSR
        if (e == cudaSuccess) e = cudaGetLastError();
        unsigned long long cnt[2] = {0, 0};
        if (e == cudaSuccess) e = cudaMemcpyAsync(cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { st = fail(ctx, MCLCAR_ECUDA, "CSR sampler: %s", cudaGetErrorString(e)); break; }
        if (accept_rate) *accept_rate = cnt[1] ? (double)cnt[0] / (double)cnt[1] : 1.0;
        if (sweeps_out) *sweeps_out = (int64_t)burn + (int64_t)E * thin;
    } while (0);
    cudaStreamSynchronize(ctx->stream);
    for (void* d : owned) pool_free(ctx, d);
    return st;
}

}  // namespace mclcar

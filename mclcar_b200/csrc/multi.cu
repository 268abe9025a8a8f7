// Single-process multi-GPU driver for the direct CAR path: what an R session (one process, one
// thread) uses on a multi-GPU box.  One context per device; the devices are driven concurrently
// by short-lived host threads (the CUDA runtime is thread-safe, each thread binds its own
// device); every context draws its own global chains (shard r of G) and reduces its own
// samples; the G partial tuples per candidate are merged on the host in shard order -- the same
// merge the NCCL path runs after its all-gather (comm.cu), so results agree with it bit for bit.
#include <thread>

#include "engine.hpp"

namespace mclcar {
int dcar_partial(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int what,
                 const int64_t* subset_idx, const int64_t* subset_ptr, int shift_mode, double* tuples, bool merge_ranks);
void* host_group_new(int G);
int dcar_finish(mclcar_ctx* ctx, const double* psi0, int P0, const double* psi, int K, int P, int what, int evar,
                const double* rho_cons, const double* tuples, double* out0, double* out1);
}  // namespace mclcar

using namespace mclcar;

static int check_multi(mclcar_ctx** ctxs, int G) {
    if (!ctxs || G < 1) return MCLCAR_EINVAL;
    for (int g = 0; g < G; ++g) {
        if (!ctxs[g]) return MCLCAR_EINVAL;
        if (ctxs[g]->host_only) return fail(ctxs[0], MCLCAR_ENODEV, "context %d is host-only: no CPU fallback", g);
        if (ctxs[g]->n != ctxs[0]->n || ctxs[g]->p != ctxs[0]->p || !ctxs[g]->has_obs)
            return fail(ctxs[0], MCLCAR_ESTATE, "context %d does not hold the same model as context 0", g);
    }
    return MCLCAR_OK;
}

/* ---- GLM and hierarchical model over several GPUs of one process.  The G contexts form an in-process group for the
 * duration of the call: every device runs the ordinary one-shot entry point on its own shard from its own host thread,
 * and where the ranks of a multi-process job exchange their partial tuples through NCCL the members of the group
 * exchange them through host memory and merge them in shard order -- every member ends with the same result. ---- */
template <class Fn>
static int on_all_devices(mclcar_ctx** ctxs, int G, bool exchange, Fn fn) {
    void* hg = exchange ? host_group_new(G) : nullptr;
    if (exchange && !hg) return fail(ctxs[0], MCLCAR_ENOMEM, "out of host memory");
    for (int g = 0; g < G; ++g) {
        mclcar_ctx_set_shard(ctxs[g], g, G);
        ctxs[g]->draw = ctxs[0]->draw;          // the shards of one draw share its Philox key
        ctxs[g]->host_group = hg;
    }
    std::vector<int> st(G, MCLCAR_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) th.emplace_back([&, g] { st[g] = fn(g); });
    for (auto& t : th) t.join();
    for (int g = 0; g < G; ++g) host_group_leave(ctxs[g]);
    for (int g = 0; g < G; ++g)
        if (st[g] != MCLCAR_OK) {
            if (g) ctxs[0]->err = ctxs[g]->err;
            return st[g];
        }
    return MCLCAR_OK;
}

static int64_t shard_count(int64_t N_total, int g, int G) { return N_total / G + (g < N_total % G ? 1 : 0); }

extern "C" {

int mclcar_multi_dcar_prep(mclcar_ctx** ctxs, int G, const double* psi0, int P, int64_t N_total,
                           const mclcar_sampler_opts* opts, mclcar_samples** out) {
    if (!psi0 || !out || N_total < G) return MCLCAR_EINVAL;
    MCL_TRY(check_multi(ctxs, G));
    std::vector<int> st(G, MCLCAR_OK);
    std::vector<std::thread> th;
    for (int g = 1; g < G; ++g) ctxs[g]->draw = ctxs[0]->draw;   // the shards of one draw share its Philox key
    for (int g = 0; g < G; ++g) {
        out[g] = nullptr;
        th.emplace_back([&, g] {
            mclcar_ctx_set_shard(ctxs[g], g, G);
            const int64_t Ng = N_total / G + (g < N_total % G ? 1 : 0);
            st[g] = mclcar_dcar_prep(ctxs[g], psi0, P, Ng, opts, &out[g]);
        });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (st[g] != MCLCAR_OK) {
            if (g) ctxs[0]->err = ctxs[g]->err;
            for (int h = 0; h < G; ++h) {
                mclcar_samples_free(out[h]);
                out[h] = nullptr;
            }
            return st[g];
        }
    return MCLCAR_OK;
}

/* what / evar / outputs as in mclcar_dcar_finish */
int mclcar_multi_dcar_run(mclcar_ctx** ctxs, mclcar_samples** samples, int G, const double* psi, int K, int P, int what,
                          int evar, const double* rho_cons, int shift_mode, double* out0, double* out1) {
    if (!samples || !psi || !out0 || K < 1) return MCLCAR_EINVAL;
    MCL_TRY(check_multi(ctxs, G));
    const int d = 2 + 2 * ctxs[0]->p;
    const int F = what == WHAT_LR ? 0 : P;
    const int64_t T = tuple_len(what, F, d);
    if (T < 0) return fail(ctxs[0], MCLCAR_EINVAL, "bad 'what'");
    std::vector<double> all((size_t)G * K * T), merged((size_t)K * T);
    std::vector<int> st(G, MCLCAR_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g)
        th.emplace_back([&, g] {
            st[g] = dcar_partial(ctxs[g], samples[g], psi, K, P, what, nullptr, nullptr, shift_mode, all.data() + (size_t)g * K * T, false);
        });
    for (auto& t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (st[g] != MCLCAR_OK) {
            if (g) ctxs[0]->err = ctxs[g]->err;
            return st[g];
        }
    MCL_TRY(merge_tuples(what, F, d, G, K, all.data(), merged.data()));
    return dcar_finish(ctxs[0], samples[0]->psi0.data(), (int)samples[0]->psi0.size(), psi, K, P, what, evar, rho_cons,
                       merged.data(), out0, out1);
}


int mclcar_multi_glm_prep(mclcar_ctx** ctxs, int G, int family, const double* psi0, int P, int64_t N_total,
                          const mclcar_sampler_opts* opts, mclcar_samples** out) {
    if (!psi0 || !out || N_total < G) return MCLCAR_EINVAL;
    MCL_TRY(check_multi(ctxs, G));
    for (int g = 0; g < G; ++g) out[g] = nullptr;
    int st = on_all_devices(ctxs, G, false, [&](int g) {
        return mclcar_glm_prep(ctxs[g], family, psi0, P, shard_count(N_total, g, G), opts, &out[g]);
    });
    if (st != MCLCAR_OK)
        for (int g = 0; g < G; ++g) { mclcar_samples_free(out[g]); out[g] = nullptr; }
    return st;
}

/* what: 0 = mcl.glm (out0 = lr[K], out1 = var[K] or NULL), 1 = mcl.grad.glm (grad[K*P], gradvar[K*P*P] when evar),
 * 2 = mcl.Hessian.glm (hess[K*P*P]) */
int mclcar_multi_glm_run(mclcar_ctx** ctxs, mclcar_samples** samples, int G, const double* psi, int K, int P, int what,
                         int evar, int shift_mode, double* out0, double* out1) {
    if (!samples || !psi || !out0 || K < 1 || what < 0 || what > 2) return MCLCAR_EINVAL;
    MCL_TRY(check_multi(ctxs, G));
    const size_t n0 = what == 0 ? (size_t)K : (what == 1 ? (size_t)K * P : (size_t)K * P * P);
    const size_t n1 = what == 0 ? (size_t)K : (size_t)K * P * P;
    std::vector<std::vector<double>> o0(G, std::vector<double>(n0)), o1(G, std::vector<double>(out1 ? n1 : 0));
    MCL_TRY(on_all_devices(ctxs, G, true, [&](int g) {
        double* a = o0[g].data();
        double* b = out1 ? o1[g].data() : nullptr;
        if (what == 0) return mclcar_glm_eval(ctxs[g], samples[g], psi, K, P, nullptr, nullptr, shift_mode, a, b);
        if (what == 1) return mclcar_glm_grad(ctxs[g], samples[g], psi, K, P, evar, shift_mode, a, b);
        return mclcar_glm_hess(ctxs[g], samples[g], psi, K, P, shift_mode, a);
    }));
    std::copy(o0[0].begin(), o0[0].end(), out0);
    if (out1) std::copy(o1[0].begin(), o1[0].end(), out1);
    return MCLCAR_OK;
}

int mclcar_multi_hcar_prep(mclcar_ctx** ctxs, int G, const double* psi0, int P, int64_t N_total,
                           const mclcar_sampler_opts* opts, mclcar_samples** out) {
    if (!psi0 || !out || N_total < G) return MCLCAR_EINVAL;
    MCL_TRY(check_multi(ctxs, G));
    for (int g = 0; g < G; ++g) out[g] = nullptr;
    int st = on_all_devices(ctxs, G, false, [&](int g) {
        return mclcar_hcar_prep(ctxs[g], psi0, P, shard_count(N_total, g, G), opts, &out[g]);
    });
    if (st != MCLCAR_OK)
        for (int g = 0; g < G; ++g) { mclcar_samples_free(out[g]); out[g] = nullptr; }
    return st;
}

int mclcar_multi_hcar_run(mclcar_ctx** ctxs, mclcar_samples** samples, int G, const double* psi, int K, int P, int what,
                          int evar, int shift_mode, double* out0, double* out1, int* n_clamped) {
    if (!samples || !psi || !out0 || K < 1 || what < 0 || what > 2) return MCLCAR_EINVAL;
    MCL_TRY(check_multi(ctxs, G));
    const size_t n0 = what == 0 ? (size_t)K : (what == 1 ? (size_t)K * P : (size_t)K * P * P);
    const size_t n1 = what == 0 ? (size_t)K : (size_t)K * P * P;
    std::vector<std::vector<double>> o0(G, std::vector<double>(n0)), o1(G, std::vector<double>(out1 ? n1 : 0));
    std::vector<int> clamped(G, 0);
    MCL_TRY(on_all_devices(ctxs, G, true, [&](int g) {
        double* a = o0[g].data();
        double* b = out1 ? o1[g].data() : nullptr;
        if (what == 0) return mclcar_hcar_eval(ctxs[g], samples[g], psi, K, P, a, b, &clamped[g]);
        if (what == 1) return mclcar_hcar_grad(ctxs[g], samples[g], psi, K, P, evar, shift_mode, a, b);
        return mclcar_hcar_hess(ctxs[g], samples[g], psi, K, P, shift_mode, a);
    }));
    std::copy(o0[0].begin(), o0[0].end(), out0);
    if (out1) std::copy(o1[0].begin(), o1[0].end(), out1);
    if (n_clamped) *n_clamped = clamped[0];
    return MCLCAR_OK;
}

}  // extern "C"

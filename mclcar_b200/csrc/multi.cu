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
                 const int64_t* subset_idx, const int64_t* subset_ptr, int shift_mode, double* tuples);
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
            st[g] = dcar_partial(ctxs[g], samples[g], psi, K, P, what, nullptr, nullptr, shift_mode, all.data() + (size_t)g * K * T);
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

}  // extern "C"

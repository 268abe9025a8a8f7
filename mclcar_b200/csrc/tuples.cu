// Partial-tuple algebra (host only, no CUDA calls): merging the tuples of several shards
// (GPUs or ranks) and the small dense helpers used by the finish stages.
//
// Tuple layouts (doubles), F = number of gradient features, V = 1 + F, d = statistics:
//   WHAT_LR   : n, m, c, S1, S2                                   (5)
//   WHAT_GRAD : n, m, centre[V], S1[V], S2[V(V+1)/2 packed upper] (2 + 2V + V(V+1)/2)
//   WHAT_HESS : n, m, Se, Sf[F], Sff[F(F+1)/2], R1[d]             (3 + F + F(F+1)/2 + d)
// where, with e_i = exp(l_i - m) and v_i = (e_i, e_i f_i):  S1 = sum (v_i - centre),
// S2 = sum (v_i - centre)(v_i - centre)',  Se = sum e_i, Sf = sum e_i f_i,
// Sff = sum e_i f_i f_i',  R1 = sum e_i q_i.
#include <cmath>

#include "engine.hpp"

namespace mclcar {

// merge G tuples (each of length T) of one candidate into `out`
static void merge_one(int what, int F, int d, int G, const double* const* in, double* out, int64_t T) {
    const int V = what == WHAT_LR ? 1 : 1 + F;
    // common shift: the largest shift among non-empty shards
    double mstar = -INFINITY;
    int first = -1;
    for (int g = 0; g < G; ++g)
        if (in[g][0] > 0) {
            if (first < 0) first = g;
            mstar = std::fmax(mstar, in[g][1]);
        }
    for (int64_t t = 0; t < T; ++t) out[t] = 0.0;
    if (first < 0) return;
    out[1] = mstar;
    if (what == WHAT_HESS) {
        const int ne = 1 + F + F * (F + 1) / 2 + d;
        for (int g = 0; g < G; ++g) {
            if (!(in[g][0] > 0)) continue;
            const double f = std::exp(in[g][1] - mstar);
            out[0] += in[g][0];
            for (int t = 0; t < ne; ++t) out[2 + t] += f * in[g][2 + t];
        }
        return;
    }
    // centred tuples: rescale to the common shift, re-centre on the first shard's centre
    const int ns2 = V * (V + 1) / 2;
    double* cstar = out + 2;
    double* S1 = out + 2 + V;
    double* S2 = out + 2 + 2 * V;
    const double f0 = std::exp(in[first][1] - mstar);
    for (int a = 0; a < V; ++a) cstar[a] = f0 * in[first][2 + a];
    std::vector<double> delta(V), s1(V);
    for (int g = 0; g < G; ++g) {
        const double n = in[g][0];
        if (!(n > 0)) continue;
        const double f = std::exp(in[g][1] - mstar);
        out[0] += n;
        for (int a = 0; a < V; ++a) {
            delta[a] = f * in[g][2 + a] - cstar[a];
            s1[a] = f * in[g][2 + V + a];
        }
        // sum (v - c*) = sum (v - c) + n (c - c*)
        for (int a = 0; a < V; ++a) S1[a] += s1[a] + n * delta[a];
        // sum (v - c*)(v - c*)' = S2 + delta s1' + s1 delta' + n delta delta'
        int s = 0;
        for (int a = 0; a < V; ++a)
            for (int b = a; b < V; ++b, ++s)
                S2[s] += f * f * in[g][2 + 2 * V + s] + delta[a] * s1[b] + s1[a] * delta[b] + n * delta[a] * delta[b];
        (void)ns2;
    }
}

int merge_tuples(int what, int F, int d, int G, int K, const double* in, double* out) {
    const int64_t T = tuple_len(what, F, d);
    if (T < 0 || G < 1 || K < 0) return MCLCAR_EINVAL;
    std::vector<const double*> ptr(G);
    for (int k = 0; k < K; ++k) {
        for (int g = 0; g < G; ++g) ptr[g] = in + ((size_t)g * K + k) * T;
        merge_one(what, F, d, G, ptr.data(), out + (size_t)k * T, T);
    }
    return MCLCAR_OK;
}

// moments recovered from a centred tuple: mean vector vbar[V] and centred co-moment C[V][V]
void centred_moments(const double* tup, int V, double* vbar, double* C) {
    const double n = tup[0];
    const double* c = tup + 2;
    const double* S1 = tup + 2 + V;
    const double* S2 = tup + 2 + 2 * V;
    for (int a = 0; a < V; ++a) vbar[a] = c[a] + S1[a] / n;
    int s = 0;
    for (int a = 0; a < V; ++a)
        for (int b = a; b < V; ++b, ++s) {
            const double v = S2[s] - S1[a] * S1[b] / n;
            C[a * V + b] = v;
            C[b * V + a] = v;
        }
}

}  // namespace mclcar

extern "C" {

int64_t mclcar_tuple_len(int what, int F, int d) { return mclcar::tuple_len(what, F, d); }

int mclcar_tuples_merge(int what, int F, int d, int G, int K, const double* tuples_in, double* tuples_out) {
    if (!tuples_in || !tuples_out) return MCLCAR_EINVAL;
    if (F < 0 || F > MCLCAR_MAX_P || d < 0) return MCLCAR_EINVAL;
    return mclcar::merge_tuples(what, F, d, G, K, tuples_in, tuples_out);
}

}  // extern "C"

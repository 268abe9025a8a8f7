// K2 (general neighbourhood matrix) -- per-sample quadratic and linear forms for a sample
// matrix stored site-major ("row = sample" of the R matrix Zy, R/postZsub.R:427: for every
// site the N samples are contiguous).  One lane = one sample, so every load of a site's
// row is a coalesced run over samples and the CSR indices are warp-uniform:
//     out[0][i] = x_i'x_i,   out[1][i] = x_i'A x_i,   out[2+l][i] = v_l'x_i
// with A a symmetric CSR matrix (W of the CAR prior, or M / Z'Z / Z'WZ of the hierarchical
// model) visited through its upper triangle, and v_l dense vectors (columns of X and WX).
// Replaces the per-sample dense products at R/mcl.dCAR.R:168-210, R/mcl.bin.R:188-196,
// R/mcl.HCAR.R:187-196.  Sites are split into chunks over blockIdx.y; the chunk partials
// are added in chunk order (deterministic).
//
// Sample-major matrices on a general graph are transposed on the device first
// (transpose_kernel) -- those are the small parity-path inputs.
#include "common.cuh"
#include "engine.hpp"

namespace mclcar {
namespace {

constexpr int kThreads = 128;

struct SiteMajorParams {
    const void* M;
    int64_t ld, N;
    int n;
    // upper triangle (j > s) of the symmetric matrix as a flat edge list sorted by s, plus the diagonal
    const int* eptr;    // [n + 1] first edge of every site
    const int* ei;      // [nedge] s
    const int* ej;      // [nedge] j
    const double* ew;   // [nedge] A_sj, null = 1
    const double* diag; // [n] A_ss, null = 0
    int L;
    const double* vec[2 * MCLCAR_MAX_COV];
    int chunk;     // sites per chunk
    double* out;   // [chunks][2+L][ldq]
    int64_t ldq;
};

template <typename T, int LT>
__global__ void __launch_bounds__(kThreads) stats_site_major_kernel(const SiteMajorParams p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    const T* M = reinterpret_cast<const T*>(p.M) + i;
    const int s0 = blockIdx.y * p.chunk;
    const int s1 = min(p.n, s0 + p.chunk);
    double xx = 0.0, xax = 0.0, lin[LT > 0 ? LT : 1];
#pragma unroll
    for (int l = 0; l < LT; ++l) lin[l] = 0.0;
    // (1) x'x, diagonal part of x'Ax and the linear forms: a streaming pass over the chunk's sites
#pragma unroll 4
    for (int s = s0; s < s1; ++s) {
        const double x = (double)ld_stream(M + (size_t)s * p.ld);
        xx = fma(x, x, xx);
        if (p.diag) xax = fma(0.5 * __ldg(p.diag + s) * x, x, xax);
#pragma unroll
        for (int l = 0; l < LT; ++l)
            if (l < p.L) lin[l] = fma(__ldg(p.vec[l] + s), x, lin[l]);
    }
    // (2) off-diagonal part: the chunk's edges, independent iterations => many gathers in flight
    const int e0 = p.eptr[s0], e1 = p.eptr[s1];
#pragma unroll 4
    for (int e = e0; e < e1; ++e) {
        const double a = (double)__ldg(M + (size_t)__ldg(p.ei + e) * p.ld);
        const double b = (double)__ldg(M + (size_t)__ldg(p.ej + e) * p.ld);
        const double w = p.ew ? __ldg(p.ew + e) : 1.0;
        xax = fma(w * a, b, xax);
    }
    double* o = p.out + (size_t)blockIdx.y * (2 + p.L) * p.ldq + i;
    o[0] = xx;
    o[p.ldq] = 2.0 * xax;
#pragma unroll
    for (int l = 0; l < LT; ++l)
        if (l < p.L) o[(size_t)(2 + l) * p.ldq] = lin[l];
}

__global__ void chunk_sum_kernel(const double* partial, double* q, int S, int d, int64_t ldq, int64_t N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int j = 0; j < d; ++j) {
        double a = 0.0;
        for (int s = 0; s < S; ++s) a += partial[((size_t)s * d + j) * ldq + i];
        q[(size_t)j * ldq + i] = a;
    }
}

template <typename TI, typename TO>
__global__ void transpose_kernel(const TI* in, TO* out, int64_t rows, int64_t cols, int64_t ld_in, int64_t ld_out) {
    __shared__ TO tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t r = r0 + k, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[k][threadIdx.x] = (TO)in[r * ld_in + c];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int64_t c = c0 + k, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[c * ld_out + r] = tile[threadIdx.x][k];
    }
}

template <typename T>
int launch_typed(mclcar_ctx* ctx, const SiteMajorParams& p, dim3 grid) {
    cudaStream_t st = ctx->stream;
    if (p.L == 0) stats_site_major_kernel<T, 0><<<grid, kThreads, 0, st>>>(p);
    else if (p.L <= 2) stats_site_major_kernel<T, 2><<<grid, kThreads, 0, st>>>(p);
    else if (p.L <= 4) stats_site_major_kernel<T, 4><<<grid, kThreads, 0, st>>>(p);
    else if (p.L <= 8) stats_site_major_kernel<T, 8><<<grid, kThreads, 0, st>>>(p);
    else stats_site_major_kernel<T, 16><<<grid, kThreads, 0, st>>>(p);
    MCL_CUDA(ctx, cudaGetLastError());
    return MCLCAR_OK;
}

}  // namespace

// generic form: any symmetric CSR matrix (host copy given), any list of dense device vectors.  The
// upper-triangle edge list is built on the host and cached on the context keyed by the CSR arrays.
// upper-triangle edge list of a symmetric CSR matrix on the device, built once per matrix and cached on the context
// keyed by the host CSR arrays
int edge_cache_get(mclcar_ctx* ctx, int n, const std::vector<int>& rowptr, const std::vector<int>& colidx,
                   const std::vector<double>& val, const EdgeCache** out) {
    EdgeCache* ec = nullptr;
    for (auto& c : ctx->edge_cache)
        if (c.key_rowptr == rowptr.data() && c.key_colidx == colidx.data() && c.n == n && c.nnz == (int)colidx.size()) ec = &c;
    if (!ec) {
        std::vector<int> eptr(n + 1, 0), ei, ej;
        std::vector<double> ew, diag(n, 0.0);
        bool any_diag = false, weighted = !val.empty();
        for (int s = 0; s < n; ++s) {
            eptr[s] = (int)ei.size();
            for (int e = rowptr[s]; e < rowptr[s + 1]; ++e) {
                const int j = colidx[e];
                const double w = weighted ? val[e] : 1.0;
                if (j == s) { diag[s] += w; any_diag = true; }
                else if (j > s) { ei.push_back(s); ej.push_back(j); ew.push_back(w); }
            }
        }
        eptr[n] = (int)ei.size();
        bool unit = true;
        for (double w : ew) unit = unit && w == 1.0;
        ctx->edge_cache.emplace_back();
        ec = &ctx->edge_cache.back();
        ec->key_rowptr = rowptr.data(); ec->key_colidx = colidx.data(); ec->n = n; ec->nnz = (int)colidx.size();
        auto up = [&](const void* h, size_t bytes, void** d) -> int {
            MCL_TRY(pool_alloc(ctx, d, bytes ? bytes : 16));
            if (bytes) MCL_CUDA(ctx, cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
            return MCLCAR_OK;
        };
        MCL_TRY(up(eptr.data(), eptr.size() * sizeof(int), (void**)&ec->d_eptr));
        MCL_TRY(up(ei.data(), ei.size() * sizeof(int), (void**)&ec->d_ei));
        MCL_TRY(up(ej.data(), ej.size() * sizeof(int), (void**)&ec->d_ej));
        if (!unit) MCL_TRY(up(ew.data(), ew.size() * sizeof(double), (void**)&ec->d_ew));
        if (any_diag) MCL_TRY(up(diag.data(), diag.size() * sizeof(double), (void**)&ec->d_diag));
        MCL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the host vectors are locals
    }
    *out = ec;
    return MCLCAR_OK;
}

int launch_quadforms_site_major(mclcar_ctx* ctx, const void* M, int dtype, int64_t N, int64_t ld, int n,
                                const std::vector<int>& rowptr, const std::vector<int>& colidx,
                                const std::vector<double>& val, int L, const double* const* d_vecs, double* d_out,
                                int64_t ldq) {
    if (L > 2 * MCLCAR_MAX_COV) return fail(ctx, MCLCAR_ELIMIT, "too many linear forms (%d)", L);
    const EdgeCache* ec = nullptr;
    MCL_TRY(edge_cache_get(ctx, n, rowptr, colidx, val, &ec));
    SiteMajorParams p{};
    p.M = M; p.ld = ld; p.N = N; p.n = n;
    p.eptr = ec->d_eptr; p.ei = ec->d_ei; p.ej = ec->d_ej; p.ew = ec->d_ew; p.diag = ec->d_diag;
    p.L = L;
    for (int l = 0; l < L; ++l) p.vec[l] = d_vecs[l];
    p.ldq = ldq;
    const int d = 2 + L;
    // split the sites so that the grid fills the machine about four times over
    const int64_t xblocks = (N + kThreads - 1) / kThreads;
    int chunks = (int)std::max<int64_t>(1, ((int64_t)8 * ctx->n_sm + xblocks - 1) / xblocks);
    chunks = std::min(chunks, std::max(1, n / 64));
    chunks = std::min(chunks, 65535);
    p.chunk = (n + chunks - 1) / chunks;
    chunks = (n + p.chunk - 1) / p.chunk;
    double* partial = d_out;
    if (chunks > 1) {
        MCL_TRY(ensure(ctx, ctx->d_partial, (size_t)chunks * d * ldq * sizeof(double)));
        partial = (double*)ctx->d_partial.p;
    }
    p.out = partial;
    dim3 grid((unsigned)xblocks, (unsigned)chunks);
    ProfScope prof(ctx, PROF_STATS, chunks > 1 ? 2 : 1);
    if (dtype == MCLCAR_F32) MCL_TRY(launch_typed<float>(ctx, p, grid));
    else MCL_TRY(launch_typed<double>(ctx, p, grid));
    if (chunks > 1) {
        chunk_sum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(partial, d_out, chunks, d, ldq, N);
        MCL_CUDA(ctx, cudaGetLastError());
    }
    return MCLCAR_OK;
}

int launch_stats_site_major(mclcar_ctx* ctx, const void* M, int dtype, int64_t N, int64_t ld, double* d_q,
                            int64_t ldq) {
    MCL_TRY(ensure_algebra(ctx));   // d_WX
    const double* vecs[2 * MCLCAR_MAX_COV];
    const int p = ctx->p;
    for (int l = 0; l < p; ++l) {
        vecs[l] = ctx->d_X + (size_t)l * ctx->n;
        vecs[p + l] = ctx->d_WX + (size_t)l * ctx->n;
    }
    return launch_quadforms_site_major(ctx, M, dtype, N, ld, ctx->n, ctx->rowptr, ctx->colidx, ctx->val, 2 * p, vecs, d_q, ldq);
}

int launch_transpose(mclcar_ctx* ctx, const void* in, int dtype, int64_t rows, int64_t cols, int64_t ld_in,
                     void* out, int64_t ld_out) {
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    if (grid.y > 65535) return fail(ctx, MCLCAR_ELIMIT, "transpose: too many rows (%lld)", (long long)rows);
    ProfScope prof(ctx, PROF_STATS, 1);
    if (dtype == MCLCAR_F32)
        transpose_kernel<float, float><<<grid, block, 0, ctx->stream>>>((const float*)in, (float*)out, rows, cols, ld_in, ld_out);
    else
        transpose_kernel<double, double><<<grid, block, 0, ctx->stream>>>((const double*)in, (double*)out, rows, cols, ld_in, ld_out);
    MCL_CUDA(ctx, cudaGetLastError());
    return MCLCAR_OK;
}

}  // namespace mclcar

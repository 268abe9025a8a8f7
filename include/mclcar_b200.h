/*
 * mclcar_b200 -- C ABI of the B200-native engine for the Monte Carlo likelihood
 * hot path of the R package shazhe/mclcar.
 *
 * The reference is pure R (DESCRIPTION:16 "NeedsCompilation: no"): there is no
 * existing FFI for this path.  The boundary is therefore the set of R functions
 * listed in SURVEY.md section 8(b); every entry point below names the R function
 * (file:line under /root/reference/mclcar_0.2-0/) whose body it replaces behind an
 * unchanged R signature.  The `.Call` glue that binds these from R is in
 * r/mclcar_b200_glue.c and documented in INTEGRATION.md.
 *
 * Conventions
 *  - plain C, no torch / R / C++ types; every function returns an int status
 *    (MCLCAR_OK == 0) and never throws or longjmps; mclcar_last_error() gives the text.
 *  - all host buffers are owned by the caller and are not retained past the call.
 *  - "sigma" is the VARIANCE, as in the reference: Q = (I - rho W)/sigma (R/mcl.dCAR.R:175).
 *  - a candidate parameter batch `psi` holds K candidates of P doubles each,
 *    candidate-contiguous: psi[k*P + j].  Matrix outputs per candidate are P x P,
 *    symmetric, stored densely: out[k*P*P + a*P + b].
 *  - direct CAR: pars = (rho, sigma, beta[p]);   P == 2 (no regression) or 2 + p.
 *    GLM:        pars = (rho, sigma, beta[p]),   P == 2 + p.
 *    HCAR:       pars = (rho, lambda, sigma.e, sigma.u, beta[p]) or, when no
 *                individual-level W was given, (lambda, sigma.e, sigma.u, beta[p]).
 *  - entry points are not re-entrant per ctx; distinct ctx objects are independent.
 *  - there is NO CPU fallback: compute entry points fail with MCLCAR_ENODEV on a
 *    host-only context (device < 0) or when no CUDA device is usable.
 */
#ifndef MCLCAR_B200_H
#define MCLCAR_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define MCLCAR_API __attribute__((visibility("default")))
#else
#define MCLCAR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mclcar_ctx mclcar_ctx;
typedef struct mclcar_samples mclcar_samples;

/* ---- status codes ------------------------------------------------------------- */
#define MCLCAR_OK 0
#define MCLCAR_EINVAL 1     /* bad argument                                           */
#define MCLCAR_ERHO 2       /* rho outside (1/min(ew), 1/max(ew)): R/CAR.simLM.R:9-11  */
#define MCLCAR_ECUDA 3      /* CUDA runtime error                                     */
#define MCLCAR_ENOMEM 4     /* device or host allocation failed                       */
#define MCLCAR_ESTATE 5     /* graph / design / observation / psi0 not set yet        */
#define MCLCAR_ENODEV 6     /* compute call on a host-only context                    */
#define MCLCAR_ENCCL 7      /* NCCL could not be loaded or reported an error          */
#define MCLCAR_ELIMIT 8     /* P, p or K beyond the compiled limits                   */

/* ---- enums (plain ints at the ABI) -------------------------------------------- */
/* memory layout of a sample matrix, in R terms */
#define MCLCAR_COL_IS_SAMPLE 0 /* simY (n x N), u.y (K x N): each sample's sites contiguous  */
#define MCLCAR_ROW_IS_SAMPLE 1 /* Zy (N x n): for each site the N samples are contiguous      */
#define MCLCAR_F64 0
#define MCLCAR_F32 1
#define MCLCAR_FAMILY_DCAR 0
#define MCLCAR_FAMILY_BINOM 1
#define MCLCAR_FAMILY_POISSON 2
#define MCLCAR_FAMILY_HCAR 3
#define MCLCAR_SHIFT_MEAN 0 /* the reference's mean shift (R/mcl.dCAR.R:26-30)              */
#define MCLCAR_SHIFT_MAX 1  /* max shift: same value where both are finite, cannot overflow */

#define MCLCAR_MAX_P 12     /* largest parameter-vector length the reduction kernels take  */
#define MCLCAR_MAX_COV 8    /* largest number of regression columns p                       */

/* ---- context -------------------------------------------------------------------- */
/* device >= 0: CUDA device ordinal.  device < 0: host-only context (graph / design /
 * observation bookkeeping, tuple merge and finish stages work; compute calls fail with
 * MCLCAR_ENODEV).  `seed` is the Philox key; in R the glue draws it from R's RNG so that
 * set.seed() still makes runs reproducible. */
MCLCAR_API int mclcar_ctx_create(int device, uint64_t seed, mclcar_ctx** out);
MCLCAR_API void mclcar_ctx_destroy(mclcar_ctx* ctx);
MCLCAR_API const char* mclcar_last_error(const mclcar_ctx* ctx); /* ctx may be NULL: last create error */
/* run all work on the given cudaStream_t (NULL: the context's own stream) */
MCLCAR_API int mclcar_ctx_set_stream(mclcar_ctx* ctx, void* cuda_stream);
MCLCAR_API int mclcar_ctx_sync(mclcar_ctx* ctx);
/* device buffers (sample matrices, chain state, scratch) are recycled by a caching allocator that keeps at most
 * half of the device memory cached; trim returns every cached free block to the driver now */
MCLCAR_API int mclcar_ctx_trim(mclcar_ctx* ctx);
/* this context is shard `rank` of `world`: samplers draw global chains rank, rank+world,
 * ... so that the union over ranks does not depend on `world`. */
MCLCAR_API int mclcar_ctx_set_shard(mclcar_ctx* ctx, int rank, int world);
/* Every sampler run (mclcar_dcar_prep, mclcar_glm_prep, mclcar_hcar_prep) uses its own Philox stream: key =
 * seed + draw * 0x9E3779B97F4A7C15 with `draw` the number of sampler runs this context has made so far --
 * as the reference draws fresh rnorm() in every mcl.prep.dCAR / postZ / sim.HCAR call (R/mcl.dCAR.R:3,
 * R/HCAR.sim.R:67).  (seed, draw) reproduces a sample set; the ranks of a sharded run must make the same
 * calls so that their counters agree.  set_draw rewinds / forwards the counter. */
MCLCAR_API int mclcar_ctx_set_draw(mclcar_ctx* ctx, uint64_t draw);
MCLCAR_API int mclcar_ctx_get_draw(const mclcar_ctx* ctx, uint64_t* draw);
/* instrumentation: kernel-launch counters per category and, while enabled, CUDA-event time
 * spans recorded on the context's stream around each group of launches.  Categories:
 * 0 torus half sweeps, 1 sample emission, 2 statistics (K2), 3 reduction (K3),
 * 4 shared-memory chain sampler, 5 GLM data terms, 6 spectrum estimate.  profile(enable) resets both. */
MCLCAR_API int mclcar_ctx_profile(mclcar_ctx* ctx, int enable);
MCLCAR_API int mclcar_ctx_profile_read(mclcar_ctx* ctx, double ms_out[8], int64_t launches_out[8]);
/* NCCL (dlopen'ed on first use, libnccl.so.2): rank 0 obtains a 128-byte unique id,
 * hands it to the other ranks by any channel, then every rank calls comm_init.  Once
 * initialised, every *_eval/_grad/_hess call all-gathers its partial tuples with ONE
 * ncclAllGather per candidate batch and merges them in rank order on every rank. */
MCLCAR_API int mclcar_nccl_unique_id(mclcar_ctx* ctx, unsigned char id_out[128]);
MCLCAR_API int mclcar_nccl_comm_init(mclcar_ctx* ctx, const unsigned char id[128]);

/* ---- model: graph, design, observations ------------------------------------------ */
/* rook-neighbour n1 x n2 torus, site (i, j) at index i*n2 + j; equals the W built by
 * CAR.simTorus (R/CAR.simLM.R:4-7) when n1 == n2.  Eigenvalues are analytic. */
MCLCAR_API int mclcar_graph_torus(mclcar_ctx* ctx, int n1, int n2);
/* general neighbourhood matrix W in CSR (0-based, symmetric, zero diagonal; val NULL =
 * binary).  Replaces `data$W` (dense n x n in the reference).  Validated on upload. */
MCLCAR_API int mclcar_graph_csr(mclcar_ctx* ctx, int n, const int* rowptr, const int* colidx, const double* val);
/* The reference hands every model its W as a matrix -- CAR.simTorus returns as.matrix(W) (R/CAR.simLM.R:19) -- so
 * the lattices of BASELINE configs 1, 2 and 5 reach this library as CSR.  mclcar_graph_csr recognises an exact
 * rook torus (binary, every row of degree 4 with neighbours (i +- 1, j), (i, j +- 1) cyclic for some n1 * n2 = n,
 * n1, n2 >= 3, site (i, j) at index i*n2 + j) and switches to the torus form: tiled red-black sampler, bulk-staged
 * statistics kernel, analytic eigenvalues (no eigen(W)).  _opts with detect_torus = 0 keeps the generic kernels. */
MCLCAR_API int mclcar_graph_csr_opts(mclcar_ctx* ctx, int n, const int* rowptr, const int* colidx, const double* val,
                                     int detect_torus);
/* 0 = no graph, 1 = torus (n1, n2 returned), 2 = general CSR; n1 / n2 may be NULL */
MCLCAR_API int mclcar_graph_kind(const mclcar_ctx* ctx, int* n1, int* n2);
/* eigenvalues of W (`data$lambda`, eigen(W) at R/mcl.bin.R:176): needed by the GLM
 * gradient / Hessian and for the rho range check on CSR graphs.  Optional. */
MCLCAR_API int mclcar_graph_set_eigenvalues(mclcar_ctx* ctx, const double* ew, int n);
/* Without eigenvalues the library estimates the spectral measure of W itself, once per graph and on the device, the
 * first time a log-determinant, a gradient / Hessian sum or the rho range needs it: stochastic Lanczos quadrature
 * (64 Rademacher probes, `steps` Lanczos steps, 0 = 80; mclcar_b200/csrc/spectrum.cu) calibrated to the exact
 * tr I, tr W, tr W^2.  It replaces the per-call eigen(W) / chol.spam of R/mcl.bin.R:22-24,50-52,176 and
 * R/mcl.HCAR.R:71-75,166-173, which cannot be formed beyond a few thousand regions.  Tolerance: 2e-3 relative on
 * log det(I - rho W), 1 % of its change between two candidates, 2e-2 on the gradient / Hessian sums; exact
 * eigenvalues (set_eigenvalues, or the analytic torus spectrum) take precedence and make every sum exact.  A graph of
 * at most 64 sites (the probe block) gets its exact eigenvalues instead (Jacobi rotations on the host). */
MCLCAR_API int mclcar_graph_estimate_spectrum(mclcar_ctx* ctx, int steps, uint64_t seed);
/* 0 = no spectrum yet, 1 = exact eigenvalues, 2 = Lanczos quadrature; *nodes (may be NULL) = number of nodes */
MCLCAR_API int mclcar_graph_spectrum_kind(const mclcar_ctx* ctx, int* nodes);
/* design matrix X = model.matrix(y ~ ., data.vec) / covX, n x p column-major; p may be 0 */
MCLCAR_API int mclcar_set_design(mclcar_ctx* ctx, const double* X, int p);
/* observed response y (n) and, for the binomial family, n.trial (n, or NULL) */
MCLCAR_API int mclcar_set_obs(mclcar_ctx* ctx, const double* y, const double* ntrial);
MCLCAR_API int mclcar_graph_n(const mclcar_ctx* ctx);
MCLCAR_API int mclcar_graph_ncolors(const mclcar_ctx* ctx);

/* ---- sample sets ------------------------------------------------------------------- */
typedef struct mclcar_sampler_opts {
    int64_t n_chains;  /* parallel chains on this context (0: chosen from N and memory)    */
    int32_t burn;      /* sweeps discarded at the start of every chain (0: default)         */
    int32_t thin;      /* sweeps between kept states (0: default)                            */
    int32_t keep;      /* 1: keep the sample matrix in HBM; 0: keep sufficient statistics only */
    int32_t dtype;     /* MCLCAR_F32 (default) or MCLCAR_F64 stored samples (chain state is fp32) */
    int32_t reserved[2];
    double omega;      /* over-relaxation factor of the Gaussian Gibbs update in [1, 2); 0: optimal
                          for the graph's spectrum (SOR); 1: plain Gibbs                        */
} mclcar_sampler_opts;

/* the reference's own sample matrix (simY / Zy / u.y) copied to the device: parity path.
 * `sites` is n for dCAR / GLM and K (districts) for HCAR. */
MCLCAR_API int mclcar_samples_from_host(mclcar_ctx* ctx, const void* M, int dtype, int layout, int64_t sites,
                             int64_t N, int64_t ld, mclcar_samples** out);
/* same, for a matrix already resident in device memory (not copied, not owned) */
MCLCAR_API int mclcar_samples_from_device(mclcar_ctx* ctx, const void* M_dev, int dtype, int layout, int64_t sites,
                               int64_t N, int64_t ld, mclcar_samples** out);
MCLCAR_API void mclcar_samples_free(mclcar_samples* s);
MCLCAR_API int64_t mclcar_samples_count(const mclcar_samples* s);
/* copy the sample matrix back as fp64 in the given layout (ld >= leading extent) */
MCLCAR_API int mclcar_samples_fetch(mclcar_ctx* ctx, mclcar_samples* s, int layout, double* out, int64_t ld);
/* sampler diagnostics: mean acceptance rate (1 for pure Gibbs) and sweeps actually run */
MCLCAR_API int mclcar_samples_diag(mclcar_ctx* ctx, mclcar_samples* s, double* accept_rate, int64_t* sweeps);
/* the schedule the sampler ran for this sample set: parallel chains, burn-in sweeps, sweeps per kept
 * state, over-relaxation factor (all 0 when the matrix came from the caller); any pointer may be NULL */
MCLCAR_API int mclcar_samples_schedule(mclcar_ctx* ctx, mclcar_samples* s, int64_t* chains, int* burn, int* thin,
                            double* omega);

/* ---- direct CAR (R/mcl.dCAR.R) --------------------------------------------------------- */
/* per-sample sufficient statistics q_i = (Y'Y, Y'WY, X'Y[p], X'WY[p]) in one pass over the
 * samples; replaces the per-sample dense products of logunnorm / D.log.unorm
 * (R/mcl.dCAR.R:168-210).  q_out: N x (2+2p) column-major, may be NULL. */
MCLCAR_API int mclcar_dcar_stats(mclcar_ctx* ctx, mclcar_samples* s, double* q_out);
/* mcl.prep.dCAR(psi, n.samples, data), R/mcl.dCAR.R:2-8: draws N samples at psi0 with the
 * chromatic Gibbs sampler and attaches t10 = h(y; psi0), t20_i = h(Y_i; psi0). */
MCLCAR_API int mclcar_dcar_prep(mclcar_ctx* ctx, const double* psi0, int P, int64_t N, const mclcar_sampler_opts* opts,
                     mclcar_samples** out);
/* Effective sample size of each of the d rows of a statistics matrix q ([d][ldq], sample index =
 * e * chains + chain, the order both samplers write): pooled autocovariances along the chains, Geyer's
 * initial positive sequence, ESS = N / tau.  Pure host code (no context, no GPU).  The reference's
 * samples are independent draws (ESS = N); this is the check that the Gibbs thinning gives the same. */
MCLCAR_API int mclcar_ess_from_stats(const double* q, int64_t N, int d, int64_t ldq, int64_t chains, double* ess_out);
/* the same for the sufficient statistics of a sample set drawn by mclcar_*_prep (d doubles out) */
MCLCAR_API int mclcar_samples_ess(mclcar_ctx* ctx, mclcar_samples* s, double* ess_out);
/* what mclcar_dcar_prep would run for (rho, N, opts) on this context's graph, without running it:
 * kind 1 = tiled red-black torus kernels, 2 = general chromatic kernel with the chains in shared
 * memory, 3 = same with the chains in global memory; chains, burn-in sweeps, sweeps per kept state,
 * over-relaxation factor.  Host logic only (also answers on a host-only context); any output may be NULL. */
MCLCAR_API int mclcar_dcar_plan(mclcar_ctx* ctx, double rho, int64_t N, const mclcar_sampler_opts* opts, int* kind,
                     int64_t* chains, int* burn, int* thin, double* omega);
/* attach the importance point to a sample set built with samples_from_*: t20 NULL means
 * "compute h(Y_i; psi0) from the statistics", otherwise the reference's own t20 (N). */
MCLCAR_API int mclcar_dcar_attach(mclcar_ctx* ctx, mclcar_samples* s, const double* psi0, int P, const double* t20);
MCLCAR_API int mclcar_dcar_t10_t20(mclcar_ctx* ctx, mclcar_samples* s, double* t10, double* t20 /* N or NULL */);
/* mcl.dCAR(pars, data, simdata, rho.cons, Evar=TRUE) for K candidates at once
 * (R/mcl.dCAR.R:11-50; batch call sites R/rsmSub.R:78-95,110).  rho_cons NULL disables the
 * guard (R: rho.cons = NULL), else candidates outside get (-1e8, 1e8).  A candidate k uses
 * the sample subset subset_idx[subset_ptr[k] .. subset_ptr[k+1]) (0-based) when subset_ptr
 * is not NULL and the range is non-empty -- the centre-point subsampling of
 * R/rsmSub.R:81-86 -- else all samples.  var may be NULL (Evar = FALSE). */
MCLCAR_API int mclcar_dcar_eval(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, const double* rho_cons,
                     const int64_t* subset_idx, const int64_t* subset_ptr, int shift_mode, double* lr,
                     double* var);
/* mc.gradlik.dCAR(pars, data, simdata, Evar), R/mcl.dCAR.R:213-283; evar in {0, 1, 2};
 * grad: K x P; gradvar: K x P x P or NULL. */
MCLCAR_API int mclcar_dcar_grad(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int evar,
                     int shift_mode, double* grad, double* gradvar);
/* mc.Hesslik.dCAR, R/mcl.dCAR.R:287-394; hess: K x P x P */
MCLCAR_API int mclcar_dcar_hess(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int shift_mode,
                     double* hess);
/* MCMLE.var, R/mcl.dCAR.R:156-165 (un-normalised, un-shifted weights); out: K x P x P */
MCLCAR_API int mclcar_dcar_mcmle_var(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, double* out);

/* ---- "next" rows of the scope: profile / GLS step and exact likelihood in batches,
 * from the p x p design constants and the spectrum of W instead of dense n x n products */
/* sigmabeta(rho, data), R/loglik.dCAR.R:72-86, for K values of rho; out: K x (1 + p) = (sigma, beta) */
MCLCAR_API int mclcar_dcar_sigmabeta(mclcar_ctx* ctx, const double* rho, int K, double* out);
/* mcl.profile.dCAR(rho, data, simdata, rho.cons, Evar), R/mcl.dCAR.R:53-96, for K values of rho at once */
MCLCAR_API int mclcar_dcar_profile_eval(mclcar_ctx* ctx, mclcar_samples* s, const double* rho, int K, const double* rho_cons,
                             int shift_mode, double* lr, double* var);
/* loglik.dCAR(pars, data, rho.cons), R/loglik.dCAR.R:7-42 (needs the eigenvalues of W), K candidates */
MCLCAR_API int mclcar_dcar_loglik(mclcar_ctx* ctx, const double* psi, int K, int P, const double* rho_cons, double* out);
/* ploglik.dCAR(rho, data), R/loglik.dCAR.R:45-69, K values of rho */
MCLCAR_API int mclcar_dcar_ploglik(mclcar_ctx* ctx, const double* rho, int K, double* out);

/* Hessian.dCAR(pars, data), R/loglik.dCAR.R:171-219: exact Hessian of the log-likelihood at `pars` (P = 2 or 2 + p) from the
 * p x p design constants, the statistics of y and the spectrum, with the reference's rho-beta block -X'r/sigma as coded
 * (:198) -- the `hessian` that summary.OptimMCL / summary.rsmMCL pass to vmle.dCAR (R/summary.OptimMCL.R:24-26).
 * out: P x P.  Host arithmetic; works on a host-only context when the eigenvalues are known. */
MCLCAR_API int mclcar_dcar_hessian_exact(mclcar_ctx* ctx, const double* pars, int P, double* out);

/* ---- Monte Carlo likelihood SURFACES and retained statistics ------------------------------
 * mcl.dCAR(c(rho_i, sigma_j, beta), data, simdata, rho.cons, Evar = TRUE) on the whole grid expand.grid(r.vec, s.vec)
 * that rsmMCL evaluates loglik.dCAR on for plot.rsmMCL (R/rsmMCL.R:26-28,95-102; drawn by plot.RSM,
 * R/rsmSub.R:315-327): rho varies fastest, lr[i + nr * j] / var[i + nr * j], i.e. the nr x ns column-major matrix of
 * `Exacts[[i]]$lrs`.  beta: pb = 0 or p regression coefficients shared by the grid.  var may be NULL. */
MCLCAR_API int mclcar_dcar_surface(mclcar_ctx* ctx, mclcar_samples* s, const double* rho, int nr, const double* sigma, int ns,
                        const double* beta, int pb, const double* rho_cons, int shift_mode, double* lr, double* var);
/* the same for mcl.glm: the only likelihood backdrop there is for a non-Gaussian family (rsmMCL stops with "No exact
 * value for non-Guassin data", R/rsmMCL.R:106); beta: the p coefficients shared by the grid */
MCLCAR_API int mclcar_glm_surface(mclcar_ctx* ctx, mclcar_samples* s, const double* rho, int nr, const double* sigma, int ns,
                       const double* beta, int shift_mode, double* lr, double* var);
/* Retained statistics.  The reference's result objects keep every iteration's sample matrix (`MC.datas`, `Sim.data`:
 * R/OptimMCL.R:80-93, R/rsmMCL.R:252-270) only so that summary() can evaluate vmle.dCAR / mc.gradlik.dCAR on it again
 * (R/summary.OptimMCL.R:24-26, R/summary.rsmMCL.R:20-22).  Every estimator of the path is a function of the N x d
 * sufficient statistics, so a result object keeps those and re-creates a sample set from them later.
 * stats_export: q_out N x d column-major, *d_out = d (q_out NULL: just ask for d); the statistics must have been
 * formed (any evaluation, mclcar_dcar_stats or mclcar_hcar_stats).  from_stats: a statistics-only sample set
 * (sites = n for the direct CAR model, K for the hierarchical one; chains = the sampler's parallel chains or 0);
 * attach the importance point afterwards with mclcar_dcar_attach / mclcar_hcar_attach. */
MCLCAR_API int mclcar_samples_stats_export(mclcar_ctx* ctx, mclcar_samples* s, int* d_out, double* q_out);
MCLCAR_API int mclcar_samples_from_stats(mclcar_ctx* ctx, const double* q, int64_t N, int d, int64_t sites, int64_t chains,
                              mclcar_samples** out);

/* ---- Binomial / Poisson GLM with latent CAR effects (R/mcl.glm.R, R/mcl.bin.R, R/mcl.pois.R) ----
 * Sample sets are the N x n matrix Zy, ROW = sample (MCLCAR_ROW_IS_SAMPLE).  The eigenvalues of
 * W must be known (torus: analytic; CSR: mclcar_graph_set_eigenvalues) -- they replace the
 * per-call chol.spam log-determinant (R/mcl.bin.R:50-52) and eigen(W) (R/mcl.bin.R:176). */
/* mcl.prep.glm + postZ (R/mcl.glm.R:5-40, R/postZ.R:2-14): N latent fields z | y at psi0 drawn by
 * the batched Metropolis-within-Gibbs sampler (site-wise independence proposals from a Gaussian approximation of
 * the full conditional, every third sweep from the prior conditional; the thinning is checked against the effective
 * sample size of z'z and z'Wz and the chains re-run when it falls short); lHZy.psi is attached implicitly. */
MCLCAR_API int mclcar_glm_prep(mclcar_ctx* ctx, int family, const double* psi0, int P, int64_t N,
                    const mclcar_sampler_opts* opts, mclcar_samples** out);
/* attach the importance point to the reference's own Zy; lHZy_psi NULL = recompute
 * log f_psi0(y, Z_i) (R/mcl.glm.R:34-37), else the reference's vector (N). */
MCLCAR_API int mclcar_glm_attach(mclcar_ctx* ctx, mclcar_samples* s, int family, const double* psi0, int P,
                      const double* lHZy_psi);
MCLCAR_API int mclcar_glm_lhzy(mclcar_ctx* ctx, mclcar_samples* s, double* out /* N */);
/* mcl.glm(pars, mcdata, family, Evar) = mcl.bin / mcl.pois (R/mcl.bin.R:33-79) for K candidates;
 * subsets as in mclcar_dcar_eval (Exp.CAR.glm centre points, R/rsmSub.R:175-181). */
MCLCAR_API int mclcar_glm_eval(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P,
                    const int64_t* subset_idx, const int64_t* subset_ptr, int shift_mode, double* lr, double* var);
/* mcl.grad.glm (R/mcl.bin.R:160-217): grad K x P, gradvar K x P x P when evar != 0 */
MCLCAR_API int mclcar_glm_grad(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int evar,
                    int shift_mode, double* grad, double* gradvar);
/* mcl.Hessian.glm (R/mcl.bin.R:222-297; Poisson beta block as coded at R/mcl.pois.R:268) */
MCLCAR_API int mclcar_glm_hess(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int shift_mode,
                    double* hess);

/* diagnostic (pure host code): the piecewise degree-8 table of L(a) = log1p(exp(-a)), a = |eta|, that the binomial data-term
 * kernel evaluates instead of exp + log1p + divide (log(1 + e^eta) = max(eta, 0) + L(|eta|), expit(|eta|) = 1 + L'),
 * with the kernel's own arithmetic: L_out[i], dL_out[i] (may be NULL).  Absolute error <= 2e-15 / 1e-12. */
MCLCAR_API int mclcar_glm_softplus_table_eval(const double* eta, int64_t n, double* L_out, double* dL_out);

/* get.beta.glm(beta0, rho.sig, mcdata, family) (R/mcl.glm.R:50-54 -> get.beta.bin, R/mcl.bin.R:83-95)
 * and the beta solve of mcl.profile.glm (R/mcl.bin.R:111-123): root of the beta block of
 * mcl.grad.glm at fixed (rho, sigma) by Newton steps on a batched forward-difference Jacobian
 * (nleqslv restated: stop when max|g| <= ftol or after maxit steps; the reference uses ftol = 1,
 * maxit = 2 / 3).  beta_out, fvec_out: p doubles (nleqslv's $x, $fvec); termcd as nleqslv's. */
MCLCAR_API int mclcar_glm_get_beta(mclcar_ctx* ctx, mclcar_samples* s, const double* rho_sig, const double* beta0,
                        int maxit, double ftol, double* beta_out, double* fvec_out, int* iters_out,
                        int* termcd_out);

/* ---- hierarchical CAR, Gaussian response (R/HCAR.sim.R, R/mcl.HCAR.R) ---------------------
 * y = X beta + Z u + e, e ~ N(0, sigma.e (I - rho W)^-1), u ~ N(0, sigma.u (I - lambda M)^-1).
 * The context's graph is the individual-level W (upload an empty CSR graph and pass has_W = 0
 * for the reference's W = NULL variant); design and observations as usual.  M (K x K,
 * symmetric) and Z (n x K) come in CSR; eig_M = data$bb, the eigenvalues of W = data$aa go
 * through mclcar_graph_set_eigenvalues.  A sample set is u.y (K x N, column = sample). */
MCLCAR_API int mclcar_hcar_set_model(mclcar_ctx* ctx, int K, const int* M_rowptr, const int* M_colidx, const double* M_val,
                          const int* Z_rowptr, const int* Z_colidx, const double* Z_val, const double* eig_M,
                          int has_W);
/* sim.HCAR(psi, data, n.samples), R/HCAR.sim.R:3-87: N draws of u | y at psi0 (chromatic Gibbs on
 * Q* = Z'Qe Z + Qu instead of one Cholesky and N backsolves) with lpsi and const.psi attached */
MCLCAR_API int mclcar_hcar_prep(mclcar_ctx* ctx, const double* psi0, int P, int64_t N, const mclcar_sampler_opts* opts,
                     mclcar_samples** out);
MCLCAR_API int mclcar_hcar_attach(mclcar_ctx* ctx, mclcar_samples* s, const double* psi0, int P, const double* lpsi);
/* the 4 + 2p statistics (r0'r0, r0'W r0, X'r0[p], X'W r0[p], u'u, u'Mu), r0 = y - Zu, of every sample: q_out N x (4 + 2p)
 * column-major, may be NULL (just form them) */
MCLCAR_API int mclcar_hcar_stats(mclcar_ctx* ctx, mclcar_samples* s, double* q_out);
MCLCAR_API int mclcar_hcar_lpsi(mclcar_ctx* ctx, mclcar_samples* s, double* lpsi /* N or NULL */, double* const_psi);
/* mcl.HCAR(pars, mcdata, data, Evar), R/mcl.HCAR.R:4-97, for K candidates.  The reference takes
 * exp() unshifted and clamps +-Inf to +-1e5 with a warning (:81-84): candidates where that would
 * have happened get +-1e5 and are counted in *n_clamped. */
MCLCAR_API int mclcar_hcar_eval(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, double* lr, double* var,
                     int* n_clamped);
/* mcl.grad.HCAR (R/mcl.HCAR.R:101-213) and mcl.Hessian.HCAR (:216-391, H.rho.beta as coded at :356) */
MCLCAR_API int mclcar_hcar_grad(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int evar,
                     int shift_mode, double* grad, double* gradvar);
MCLCAR_API int mclcar_hcar_hess(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int shift_mode,
                     double* hess);

/* ---- one process, several GPUs (what an R session uses): one context per device, each holding the
 * same graph / design / observations; the devices are driven concurrently by host threads and the
 * partial tuples are merged on the host in shard order (same result as the NCCL path). */
/* mcl.prep.dCAR with the N_total samples split over G contexts (shard g draws global chains g, g+G, ...) */
MCLCAR_API int mclcar_multi_dcar_prep(mclcar_ctx** ctxs, int G, const double* psi0, int P, int64_t N_total,
                           const mclcar_sampler_opts* opts, mclcar_samples** out /* G handles */);
/* mcl.dCAR / mc.gradlik.dCAR / mc.Hesslik.dCAR / MCMLE.var over all shards; what, evar, out0, out1 as in
 * mclcar_dcar_finish */
MCLCAR_API int mclcar_multi_dcar_run(mclcar_ctx** ctxs, mclcar_samples** samples, int G, const double* psi, int K, int P, int what,
                          int evar, const double* rho_cons, int shift_mode, double* out0, double* out1);

/* the same for the GLM and the hierarchical model: mcl.prep.glm / sim.HCAR with N_total samples split over the G
 * contexts, and mcl.glm / mcl.grad.glm / mcl.Hessian.glm (mcl.HCAR / ...) over all shards -- what: 0 = log-ratio
 * (out0 = lr[K], out1 = var[K] or NULL), 1 = gradient (out0 = grad[K*P], out1 = gradvar[K*P*P] or NULL), 2 = Hessian
 * (out0 = hess[K*P*P]).  Every context must hold the same model (graph, design, observations; hcar_set_model). */
MCLCAR_API int mclcar_multi_glm_prep(mclcar_ctx** ctxs, int G, int family, const double* psi0, int P, int64_t N_total,
                          const mclcar_sampler_opts* opts, mclcar_samples** out /* G handles */);
MCLCAR_API int mclcar_multi_glm_run(mclcar_ctx** ctxs, mclcar_samples** samples, int G, const double* psi, int K, int P, int what,
                         int evar, int shift_mode, double* out0, double* out1);
MCLCAR_API int mclcar_multi_hcar_prep(mclcar_ctx** ctxs, int G, const double* psi0, int P, int64_t N_total,
                           const mclcar_sampler_opts* opts, mclcar_samples** out /* G handles */);
MCLCAR_API int mclcar_multi_hcar_run(mclcar_ctx** ctxs, mclcar_samples** samples, int G, const double* psi, int K, int P, int what,
                          int evar, int shift_mode, double* out0, double* out1, int* n_clamped);

/* ---- staged evaluation (what the one-shot calls above are made of) ------------------------
 * stage 1, device: per-candidate partial tuples of this context's samples, to host memory;
 * stage 2, host only: merge the tuples of G shards in shard order (mclcar_tuples_merge);
 * stage 3, host only: tuples -> the reference's outputs (mclcar_dcar_finish).
 * `what`: 0 = lr/var, 1 = gradient (+ variance forms), 2 = Hessian.  A tuple is
 * mclcar_tuple_len(what, F, d) doubles with F = P gradient features and d = 2 + 2p
 * statistics for the direct CAR model (layouts: mclcar_b200/csrc/tuples.cu). */
MCLCAR_API int64_t mclcar_tuple_len(int what, int F, int d);
MCLCAR_API int mclcar_dcar_partial(mclcar_ctx* ctx, mclcar_samples* s, const double* psi, int K, int P, int what,
                        const int64_t* subset_idx, const int64_t* subset_ptr, int shift_mode, double* tuples);
/* tuples_in: [G][K][T], tuples_out: [K][T] */
MCLCAR_API int mclcar_tuples_merge(int what, int F, int d, int G, int K, const double* tuples_in, double* tuples_out);
/* what = 0: out0 = lr[K], out1 = var[K] or NULL.  what = 1: out0 = grad[K*P], out1 =
 * gradvar[K*P*P] or NULL with evar 1 | 2 as in mc.gradlik.dCAR, 3 = MCMLE.var.
 * what = 2: out0 = hess[K*P*P].  Works on a host-only context. */
MCLCAR_API int mclcar_dcar_finish(mclcar_ctx* ctx, const double* psi0, int P0, const double* psi, int K, int P, int what,
                       int evar, const double* rho_cons, const double* tuples, double* out0, double* out1);

#ifdef __cplusplus
}
#endif
#endif /* MCLCAR_B200_H */

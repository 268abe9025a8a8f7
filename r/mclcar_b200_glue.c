/*
 * .Call glue between the R package mclcar and libmclcar_b200.so (include/mclcar_b200.h).
 *
 * NOT COMPILED IN THIS REPOSITORY: R (Rinternals.h, libR) is not installed in the build image.
 * The file is deliberately mechanical -- argument unpacking, one C-ABI call, result packing --
 * and every call sequence below is exercised through the same C ABI by mclcar_b200/api.py
 * (ctypes) in tests/.  Build inside the R package with
 *     PKG_CPPFLAGS = -I<repo>/include      PKG_LIBS = -L<repo>/mclcar_b200/lib -lmclcar_b200
 * and register with useDynLib(mclcar, .registration = TRUE) in NAMESPACE.
 *
 * Rules followed: R errors are raised only AFTER the C call returned
 * (Rf_error longjmps); device objects live in external pointers with finalizers; the seed is
 * drawn from R's RNG so that set.seed() keeps runs reproducible; R is single threaded and
 * the library is called from that thread only.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Random.h>
#include <stdint.h>
#include <string.h>
#include "mclcar_b200.h"

static void ctx_finalizer(SEXP p) {
    mclcar_ctx* c = (mclcar_ctx*)R_ExternalPtrAddr(p);
    if (c) { mclcar_ctx_destroy(c); R_ClearExternalPtr(p); }
}
static void samples_finalizer(SEXP p) {
    mclcar_samples* s = (mclcar_samples*)R_ExternalPtrAddr(p);
    if (s) { mclcar_samples_free(s); R_ClearExternalPtr(p); }
}
static mclcar_ctx* ctx_of(SEXP p) {
    mclcar_ctx* c = (mclcar_ctx*)R_ExternalPtrAddr(p);
    if (!c) Rf_error("mclcar_b200: context already released");
    return c;
}
static mclcar_samples* samples_of(SEXP p) {
    mclcar_samples* s = (mclcar_samples*)R_ExternalPtrAddr(p);
    if (!s) Rf_error("mclcar_b200: sample set already released");
    return s;
}
#define CHECK(ctx, call)                                            \
    do {                                                            \
        int st_ = (call);                                           \
        if (st_ != MCLCAR_OK) Rf_error("%s", mclcar_last_error(ctx)); \
    } while (0)

/* ctx <- .Call(C_mclb_ctx, device): seed from R's RNG stream */
SEXP C_mclb_ctx(SEXP device) {
    GetRNGstate();
    uint64_t seed = ((uint64_t)(unif_rand() * 4294967296.0) << 32) | (uint64_t)(unif_rand() * 4294967296.0);
    PutRNGstate();
    mclcar_ctx* c = NULL;
    if (mclcar_ctx_create(Rf_asInteger(device), seed, &c) != MCLCAR_OK) Rf_error("%s", mclcar_last_error(NULL));
    SEXP p = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* graph: torus (n1, n2) or the dgCMatrix / spam slots of a symmetric W converted to 0-based CSR in R */
SEXP C_mclb_graph_torus(SEXP ctx, SEXP n1, SEXP n2) {
    mclcar_ctx* c = ctx_of(ctx);
    CHECK(c, mclcar_graph_torus(c, Rf_asInteger(n1), Rf_asInteger(n2)));
    return R_NilValue;
}
SEXP C_mclb_graph_csr(SEXP ctx, SEXP rowptr, SEXP colidx, SEXP val, SEXP eig) {
    mclcar_ctx* c = ctx_of(ctx);
    int n = LENGTH(rowptr) - 1;
    CHECK(c, mclcar_graph_csr(c, n, INTEGER(rowptr), INTEGER(colidx), Rf_isNull(val) ? NULL : REAL(val)));
    if (!Rf_isNull(eig)) CHECK(c, mclcar_graph_set_eigenvalues(c, REAL(eig), LENGTH(eig)));
    return R_NilValue;
}
/* X = model.matrix(y ~ ., data.vec) (n x p, column-major as R stores it), y, n.trial */
SEXP C_mclb_model(SEXP ctx, SEXP X, SEXP y, SEXP ntrial) {
    mclcar_ctx* c = ctx_of(ctx);
    int p = Rf_isNull(X) ? 0 : Rf_ncols(X);
    CHECK(c, mclcar_set_design(c, p ? REAL(X) : NULL, p));
    CHECK(c, mclcar_set_obs(c, REAL(y), Rf_isNull(ntrial) ? NULL : REAL(ntrial)));
    return R_NilValue;
}

static void fill_opts(SEXP control, mclcar_sampler_opts* o) { /* list(n.chains, burn, thin, keep, dtype, omega) */
    o->n_chains = (int64_t)Rf_asReal(VECTOR_ELT(control, 0));
    o->burn = Rf_asInteger(VECTOR_ELT(control, 1));
    o->thin = Rf_asInteger(VECTOR_ELT(control, 2));
    o->keep = Rf_asLogical(VECTOR_ELT(control, 3));
    o->dtype = Rf_asInteger(VECTOR_ELT(control, 4));
    o->omega = Rf_asReal(VECTOR_ELT(control, 5));
}
static SEXP wrap_samples(mclcar_samples* s) {
    SEXP p = PROTECT(R_MakeExternalPtr(s, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, samples_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* mcl.prep.dCAR(psi, n.samples, data) -> list(simY = <handle>, t10, t20)  [R/mcl.dCAR.R:2-8] */
SEXP C_mclb_prep_dcar(SEXP ctx, SEXP psi, SEXP nsamp, SEXP control) {
    mclcar_ctx* c = ctx_of(ctx);
    mclcar_sampler_opts o = {0};
    fill_opts(control, &o);
    mclcar_samples* s = NULL;
    CHECK(c, mclcar_dcar_prep(c, REAL(psi), LENGTH(psi), (int64_t)Rf_asReal(nsamp), &o, &s));
    return wrap_samples(s);
}
/* the reference's own simY (n x N) and optionally t20: parity path */
SEXP C_mclb_samples_dcar(SEXP ctx, SEXP simY, SEXP psi, SEXP t20) {
    mclcar_ctx* c = ctx_of(ctx);
    mclcar_samples* s = NULL;
    CHECK(c, mclcar_samples_from_host(c, REAL(simY), MCLCAR_F64, MCLCAR_COL_IS_SAMPLE, Rf_nrows(simY), Rf_ncols(simY),
                                      Rf_nrows(simY), &s));
    int st = mclcar_dcar_attach(c, s, REAL(psi), LENGTH(psi), Rf_isNull(t20) ? NULL : REAL(t20));
    if (st != MCLCAR_OK) { mclcar_samples_free(s); Rf_error("%s", mclcar_last_error(c)); }
    return wrap_samples(s);
}
/* mcl.dCAR over the ROWS of `pars` (K x P): returns K x 2 (mc.lr, v.lr)  [R/mcl.dCAR.R:11-50] */
SEXP C_mclb_mcl_dcar(SEXP ctx, SEXP samples, SEXP pars, SEXP rho_cons, SEXP sub_idx, SEXP sub_ptr) {
    mclcar_ctx* c = ctx_of(ctx);
    int K = Rf_nrows(pars), P = Rf_ncols(pars);
    double* psi = (double*)R_alloc((size_t)K * P, sizeof(double)); /* candidate-contiguous copy */
    for (int k = 0; k < K; ++k) for (int j = 0; j < P; ++j) psi[k * P + j] = REAL(pars)[k + (size_t)j * K];
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, K, 2));
    int64_t *idx = NULL, *ptr = NULL;
    if (!Rf_isNull(sub_ptr)) { /* 0-based indices as doubles from R */
        idx = (int64_t*)R_alloc(LENGTH(sub_idx) + 1, sizeof(int64_t));
        ptr = (int64_t*)R_alloc(LENGTH(sub_ptr), sizeof(int64_t));
        for (int i = 0; i < LENGTH(sub_idx); ++i) idx[i] = (int64_t)REAL(sub_idx)[i];
        for (int i = 0; i < LENGTH(sub_ptr); ++i) ptr[i] = (int64_t)REAL(sub_ptr)[i];
    }
    int st = mclcar_dcar_eval(c, samples_of(samples), psi, K, P, Rf_isNull(rho_cons) ? NULL : REAL(rho_cons), idx, ptr,
                              MCLCAR_SHIFT_MEAN, REAL(out), REAL(out) + K);
    UNPROTECT(1);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return out;
}
/* mc.gradlik.dCAR(pars, data, simdata, Evar) -> list(grad, grad.var)  [R/mcl.dCAR.R:213-283] */
SEXP C_mclb_grad_dcar(SEXP ctx, SEXP samples, SEXP pars, SEXP evar) {
    mclcar_ctx* c = ctx_of(ctx);
    int P = LENGTH(pars), ev = Rf_asInteger(evar);
    SEXP g = PROTECT(Rf_allocVector(REALSXP, P));
    SEXP v = PROTECT(Rf_allocMatrix(REALSXP, P, P));
    int st = mclcar_dcar_grad(c, samples_of(samples), REAL(pars), 1, P, ev, MCLCAR_SHIFT_MEAN, REAL(g), ev ? REAL(v) : NULL);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, g);
    SET_VECTOR_ELT(out, 1, ev ? v : R_NilValue);
    UNPROTECT(3);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return out;
}
/* mc.Hesslik.dCAR / MCMLE.var  [R/mcl.dCAR.R:287-394, 156-165] */
SEXP C_mclb_hess_dcar(SEXP ctx, SEXP samples, SEXP pars, SEXP which) {
    mclcar_ctx* c = ctx_of(ctx);
    int P = LENGTH(pars);
    SEXP H = PROTECT(Rf_allocMatrix(REALSXP, P, P));
    int st = Rf_asInteger(which) == 0 ? mclcar_dcar_hess(c, samples_of(samples), REAL(pars), 1, P, MCLCAR_SHIFT_MEAN, REAL(H))
                                      : mclcar_dcar_mcmle_var(c, samples_of(samples), REAL(pars), 1, P, REAL(H));
    UNPROTECT(1);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return H;
}
/* simdata$simY materialised (n x N) */
SEXP C_mclb_fetch(SEXP ctx, SEXP samples, SEXP nrow, SEXP layout) {
    mclcar_ctx* c = ctx_of(ctx);
    mclcar_samples* s = samples_of(samples);
    int64_t N = mclcar_samples_count(s);
    int rows = Rf_asInteger(nrow), lay = Rf_asInteger(layout);
    SEXP out = PROTECT(lay == MCLCAR_COL_IS_SAMPLE ? Rf_allocMatrix(REALSXP, rows, (int)N) : Rf_allocMatrix(REALSXP, (int)N, rows));
    int st = mclcar_samples_fetch(c, s, lay, REAL(out), lay == MCLCAR_COL_IS_SAMPLE ? rows : N);
    UNPROTECT(1);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return out;
}

/* sampler diagnostics of a sample set: c(accept, sweeps, chains, burn, thin, omega) */
SEXP C_mclb_samples_diag(SEXP ctx, SEXP samples) {
    mclcar_ctx* c = ctx_of(ctx);
    mclcar_samples* s = samples_of(samples);
    double acc = 0.0, omega = 0.0;
    int64_t sweeps = 0, chains = 0;
    int burn = 0, thin = 0;
    CHECK(c, mclcar_samples_diag(c, s, &acc, &sweeps));
    CHECK(c, mclcar_samples_schedule(c, s, &chains, &burn, &thin, &omega));
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 6));
    REAL(out)[0] = acc; REAL(out)[1] = (double)sweeps; REAL(out)[2] = (double)chains;
    REAL(out)[3] = burn; REAL(out)[4] = thin; REAL(out)[5] = omega;
    UNPROTECT(1);
    return out;
}

/* GLM: mcl.prep.glm / mcl.glm / mcl.grad.glm / mcl.Hessian.glm  [R/mcl.glm.R, R/mcl.bin.R, R/mcl.pois.R] */
SEXP C_mclb_prep_glm(SEXP ctx, SEXP family, SEXP psi, SEXP nsamp, SEXP control) {
    mclcar_ctx* c = ctx_of(ctx);
    mclcar_sampler_opts o = {0};
    fill_opts(control, &o);
    mclcar_samples* s = NULL;
    CHECK(c, mclcar_glm_prep(c, Rf_asInteger(family), REAL(psi), LENGTH(psi), (int64_t)Rf_asReal(nsamp), &o, &s));
    return wrap_samples(s);
}
SEXP C_mclb_samples_glm(SEXP ctx, SEXP Zy, SEXP family, SEXP psi, SEXP lHZy) { /* Zy: N x n, row = sample */
    mclcar_ctx* c = ctx_of(ctx);
    mclcar_samples* s = NULL;
    CHECK(c, mclcar_samples_from_host(c, REAL(Zy), MCLCAR_F64, MCLCAR_ROW_IS_SAMPLE, Rf_ncols(Zy), Rf_nrows(Zy), Rf_nrows(Zy), &s));
    int st = mclcar_glm_attach(c, s, Rf_asInteger(family), REAL(psi), LENGTH(psi), Rf_isNull(lHZy) ? NULL : REAL(lHZy));
    if (st != MCLCAR_OK) { mclcar_samples_free(s); Rf_error("%s", mclcar_last_error(c)); }
    return wrap_samples(s);
}
/* a new importance point for an existing latent sample set after the design of its context was replaced: the
 * Poisson-response hierarchical extension (r/overrides.R, mcl.HCAR.pois) evaluates the district-level draws against
 * designs whose columns are per-candidate offsets */
SEXP C_mclb_attach_glm(SEXP ctx, SEXP samples, SEXP family, SEXP psi) {
    mclcar_ctx* c = ctx_of(ctx);
    CHECK(c, mclcar_glm_attach(c, samples_of(samples), Rf_asInteger(family), REAL(psi), LENGTH(psi), NULL));
    return R_NilValue;
}
/* pars: K x P -> K x 2; sub_idx / sub_ptr (doubles, 0-based, or NULL): sample subsets of the centre
 * points of Exp.CAR.glm [R/rsmSub.R:175-181] */
SEXP C_mclb_mcl_glm(SEXP ctx, SEXP samples, SEXP pars, SEXP sub_idx, SEXP sub_ptr) {
    mclcar_ctx* c = ctx_of(ctx);
    int K = Rf_nrows(pars), P = Rf_ncols(pars);
    double* psi = (double*)R_alloc((size_t)K * P, sizeof(double));
    for (int k = 0; k < K; ++k) for (int j = 0; j < P; ++j) psi[k * P + j] = REAL(pars)[k + (size_t)j * K];
    int64_t *idx = NULL, *ptr = NULL;
    if (!Rf_isNull(sub_ptr)) {
        ptr = (int64_t*)R_alloc((size_t)K + 1, sizeof(int64_t));
        for (int k = 0; k <= K; ++k) ptr[k] = (int64_t)REAL(sub_ptr)[k];
        idx = (int64_t*)R_alloc((size_t)LENGTH(sub_idx) + 1, sizeof(int64_t));
        for (R_xlen_t t = 0; t < XLENGTH(sub_idx); ++t) idx[t] = (int64_t)REAL(sub_idx)[t];
    }
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, K, 2));
    int st = mclcar_glm_eval(c, samples_of(samples), psi, K, P, idx, ptr, MCLCAR_SHIFT_MEAN, REAL(out), REAL(out) + K);
    UNPROTECT(1);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return out;
}
SEXP C_mclb_grad_glm(SEXP ctx, SEXP samples, SEXP pars, SEXP evar) {
    mclcar_ctx* c = ctx_of(ctx);
    int P = LENGTH(pars), ev = Rf_asLogical(evar);
    SEXP g = PROTECT(Rf_allocVector(REALSXP, P));
    SEXP v = PROTECT(Rf_allocMatrix(REALSXP, P, P));
    int st = mclcar_glm_grad(c, samples_of(samples), REAL(pars), 1, P, ev, MCLCAR_SHIFT_MEAN, REAL(g), ev ? REAL(v) : NULL);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, g);
    SET_VECTOR_ELT(out, 1, ev ? v : R_NilValue);
    UNPROTECT(3);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return out;
}
SEXP C_mclb_hess_glm(SEXP ctx, SEXP samples, SEXP pars) {
    mclcar_ctx* c = ctx_of(ctx);
    int P = LENGTH(pars);
    SEXP H = PROTECT(Rf_allocMatrix(REALSXP, P, P));
    int st = mclcar_glm_hess(c, samples_of(samples), REAL(pars), 1, P, MCLCAR_SHIFT_MEAN, REAL(H));
    UNPROTECT(1);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return H;
}

/* get.beta.glm / the beta solve of mcl.profile.glm [R/mcl.bin.R:83-95, 111-123] -> list(x, fvec, termcd, iter) */
SEXP C_mclb_get_beta_glm(SEXP ctx, SEXP samples, SEXP rho_sig, SEXP beta0, SEXP maxit, SEXP ftol) {
    mclcar_ctx* c = ctx_of(ctx);
    int p = LENGTH(beta0), it = 0, tc = 0;
    SEXP x = PROTECT(Rf_allocVector(REALSXP, p));
    SEXP f = PROTECT(Rf_allocVector(REALSXP, p));
    int st = mclcar_glm_get_beta(c, samples_of(samples), REAL(rho_sig), REAL(beta0), Rf_asInteger(maxit), Rf_asReal(ftol),
                                 REAL(x), REAL(f), &it, &tc);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
    SET_VECTOR_ELT(out, 0, x);
    SET_VECTOR_ELT(out, 1, f);
    SET_VECTOR_ELT(out, 2, Rf_ScalarInteger(tc));
    SET_VECTOR_ELT(out, 3, Rf_ScalarInteger(it));
    UNPROTECT(3);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return out;
}

/* HCAR: sim.HCAR / mcl.HCAR / mcl.grad.HCAR / mcl.Hessian.HCAR  [R/HCAR.sim.R, R/mcl.HCAR.R] */
SEXP C_mclb_hcar_model(SEXP ctx, SEXP Mrp, SEXP Mci, SEXP Mv, SEXP Zrp, SEXP Zci, SEXP Zv, SEXP bb, SEXP hasW) {
    mclcar_ctx* c = ctx_of(ctx);
    CHECK(c, mclcar_hcar_set_model(c, LENGTH(Mrp) - 1, INTEGER(Mrp), INTEGER(Mci), REAL(Mv), INTEGER(Zrp), INTEGER(Zci),
                                   REAL(Zv), REAL(bb), Rf_asLogical(hasW)));
    return R_NilValue;
}
SEXP C_mclb_sim_hcar(SEXP ctx, SEXP psi, SEXP nsamp, SEXP control) {
    mclcar_ctx* c = ctx_of(ctx);
    mclcar_sampler_opts o = {0};
    fill_opts(control, &o);
    mclcar_samples* s = NULL;
    CHECK(c, mclcar_hcar_prep(c, REAL(psi), LENGTH(psi), (int64_t)Rf_asReal(nsamp), &o, &s));
    return wrap_samples(s);
}
SEXP C_mclb_mcl_hcar(SEXP ctx, SEXP samples, SEXP pars) { /* pars: K x P -> list(K x 2, n.clamped) */
    mclcar_ctx* c = ctx_of(ctx);
    int K = Rf_nrows(pars), P = Rf_ncols(pars), ncl = 0;
    double* psi = (double*)R_alloc((size_t)K * P, sizeof(double));
    for (int k = 0; k < K; ++k) for (int j = 0; j < P; ++j) psi[k * P + j] = REAL(pars)[k + (size_t)j * K];
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, K, 2));
    int st = mclcar_hcar_eval(c, samples_of(samples), psi, K, P, REAL(out), REAL(out) + K, &ncl);
    UNPROTECT(1);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    if (ncl) Rf_warning("Design area too large: Inf produced! Choose a parameter closer to psi."); /* R/mcl.HCAR.R:82 */
    return out;
}
SEXP C_mclb_grad_hcar(SEXP ctx, SEXP samples, SEXP pars, SEXP evar) { /* -> list(grad.lr, grad.var | NULL)  [R/mcl.HCAR.R:101-213] */
    mclcar_ctx* c = ctx_of(ctx);
    int P = LENGTH(pars), ev = Rf_asLogical(evar);
    SEXP g = PROTECT(Rf_allocVector(REALSXP, P));
    SEXP v = PROTECT(Rf_allocMatrix(REALSXP, P, P));
    int st = mclcar_hcar_grad(c, samples_of(samples), REAL(pars), 1, P, ev, MCLCAR_SHIFT_MEAN, REAL(g), ev ? REAL(v) : NULL);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, g);
    SET_VECTOR_ELT(out, 1, ev ? v : R_NilValue);
    UNPROTECT(3);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return out;
}
SEXP C_mclb_hess_hcar(SEXP ctx, SEXP samples, SEXP pars) { /* P x P  [R/mcl.HCAR.R:216-391] */
    mclcar_ctx* c = ctx_of(ctx);
    int P = LENGTH(pars);
    SEXP H = PROTECT(Rf_allocMatrix(REALSXP, P, P));
    int st = mclcar_hcar_hess(c, samples_of(samples), REAL(pars), 1, P, MCLCAR_SHIFT_MEAN, REAL(H));
    UNPROTECT(1);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return H;
}


/* ---- graph bookkeeping: what the library made of an uploaded W ------------------------------------------------- */
/* c(kind, n1, n2): kind 1 = rook torus recognised in the CSR upload (tiled sampler, analytic spectrum), 2 = general */
SEXP C_mclb_graph_kind(SEXP ctx) {
    mclcar_ctx* c = ctx_of(ctx);
    int n1 = 0, n2 = 0;
    int kind = mclcar_graph_kind(c, &n1, &n2);
    SEXP out = PROTECT(Rf_allocVector(INTSXP, 3));
    INTEGER(out)[0] = kind; INTEGER(out)[1] = n1; INTEGER(out)[2] = n2;
    UNPROTECT(1);
    return out;
}
SEXP C_mclb_set_eigenvalues(SEXP ctx, SEXP eig) {
    mclcar_ctx* c = ctx_of(ctx);
    CHECK(c, mclcar_graph_set_eigenvalues(c, REAL(eig), LENGTH(eig)));
    return R_NilValue;
}
/* CAR.simTorus(n1, n2, rho, prec)$X (R/CAR.simLM.R:2-22): one draw of N(0, (prec (I - rho W))^-1) on the n1 x n2 torus by
 * the tiled red-black sampler instead of a sparse Cholesky of the n x n precision */
SEXP C_mclb_sim_torus(SEXP n1, SEXP n2, SEXP rho, SEXP prec) {
    const int a = Rf_asInteger(n1), b = Rf_asInteger(n2);
    GetRNGstate();
    uint64_t seed = ((uint64_t)(unif_rand() * 4294967296.0) << 32) | (uint64_t)(unif_rand() * 4294967296.0);
    PutRNGstate();
    mclcar_ctx* c = NULL;
    if (mclcar_ctx_create(0, seed, &c) != MCLCAR_OK) Rf_error("%s", mclcar_last_error(NULL));
    SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)a * b));
    double* zero = (double*)R_alloc((size_t)a * b, sizeof(double));
    for (R_xlen_t i = 0; i < (R_xlen_t)a * b; ++i) zero[i] = 0.0;
    double psi[2] = {Rf_asReal(rho), 1.0 / Rf_asReal(prec)};
    mclcar_sampler_opts o = {0};
    o.keep = 1; o.dtype = MCLCAR_F64;
    mclcar_samples* s = NULL;
    int st = mclcar_graph_torus(c, a, b);
    if (st == MCLCAR_OK) st = mclcar_set_design(c, NULL, 0);
    if (st == MCLCAR_OK) st = mclcar_set_obs(c, zero, NULL);
    if (st == MCLCAR_OK) st = mclcar_dcar_prep(c, psi, 2, 1, &o, &s);
    if (st == MCLCAR_OK) st = mclcar_samples_fetch(c, s, MCLCAR_COL_IS_SAMPLE, REAL(out), (int64_t)a * b);
    char msg[512];
    if (st != MCLCAR_OK) { strncpy(msg, mclcar_last_error(c), sizeof msg - 1); msg[sizeof msg - 1] = 0; }
    mclcar_samples_free(s);
    mclcar_ctx_destroy(c);
    UNPROTECT(1);
    if (st != MCLCAR_OK) Rf_error("%s", msg);     /* "rho should be within ..." as R/CAR.simLM.R:11 */
    return out;
}

/* ---- profile / exact pieces the summaries and plots use ----------------------------------------- */
/* sigmabeta(rho, data) for a vector of rho: K x (1 + p)  [R/loglik.dCAR.R:72-86] */
SEXP C_mclb_sigmabeta(SEXP ctx, SEXP rho, SEXP p) {
    mclcar_ctx* c = ctx_of(ctx);
    int K = LENGTH(rho), w = 1 + Rf_asInteger(p);
    double* tmp = (double*)R_alloc((size_t)K * w, sizeof(double));
    int st = mclcar_dcar_sigmabeta(c, REAL(rho), K, tmp);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, K, w));
    for (int k = 0; k < K; ++k) for (int j = 0; j < w; ++j) REAL(out)[k + (size_t)j * K] = tmp[(size_t)k * w + j];
    UNPROTECT(1);
    return out;
}
/* loglik.dCAR over the rows of pars (K x P)  [R/loglik.dCAR.R:7-42]: the exact backdrop of plot.rsmMCL in one call */
SEXP C_mclb_loglik_dcar(SEXP ctx, SEXP pars, SEXP rho_cons) {
    mclcar_ctx* c = ctx_of(ctx);
    int K = Rf_nrows(pars), P = Rf_ncols(pars);
    double* psi = (double*)R_alloc((size_t)K * P, sizeof(double));
    for (int k = 0; k < K; ++k) for (int j = 0; j < P; ++j) psi[(size_t)k * P + j] = REAL(pars)[k + (size_t)j * K];
    SEXP out = PROTECT(Rf_allocVector(REALSXP, K));
    int st = mclcar_dcar_loglik(c, psi, K, P, Rf_isNull(rho_cons) ? NULL : REAL(rho_cons), REAL(out));
    UNPROTECT(1);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return out;
}
/* Hessian.dCAR(pars, data): P x P  [R/loglik.dCAR.R:171-219] */
SEXP C_mclb_hessian_dcar(SEXP ctx, SEXP pars) {
    mclcar_ctx* c = ctx_of(ctx);
    int P = LENGTH(pars);
    SEXP H = PROTECT(Rf_allocMatrix(REALSXP, P, P));
    int st = mclcar_dcar_hessian_exact(c, REAL(pars), P, REAL(H));
    UNPROTECT(1);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return H;
}
/* Monte Carlo likelihood surface on expand.grid(r.vec, s.vec): list(lrs, vars), each length(r.vec) x length(s.vec)
 * -- the shape of Exacts[[i]]$lrs (R/rsmMCL.R:95-102); family 0 = direct CAR (rho_cons may be NULL), 1/2 = GLM */
SEXP C_mclb_surface(SEXP ctx, SEXP samples, SEXP family, SEXP rvec, SEXP svec, SEXP beta, SEXP rho_cons) {
    mclcar_ctx* c = ctx_of(ctx);
    int nr = LENGTH(rvec), ns = LENGTH(svec), pb = Rf_isNull(beta) ? 0 : LENGTH(beta);
    SEXP lr = PROTECT(Rf_allocMatrix(REALSXP, nr, ns));
    SEXP vr = PROTECT(Rf_allocMatrix(REALSXP, nr, ns));
    int st;
    if (Rf_asInteger(family) == MCLCAR_FAMILY_DCAR)
        st = mclcar_dcar_surface(c, samples_of(samples), REAL(rvec), nr, REAL(svec), ns, pb ? REAL(beta) : NULL, pb,
                                 Rf_isNull(rho_cons) ? NULL : REAL(rho_cons), MCLCAR_SHIFT_MEAN, REAL(lr), REAL(vr));
    else
        st = mclcar_glm_surface(c, samples_of(samples), REAL(rvec), nr, REAL(svec), ns, REAL(beta), MCLCAR_SHIFT_MEAN,
                                REAL(lr), REAL(vr));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, lr);
    SET_VECTOR_ELT(out, 1, vr);
    UNPROTECT(3);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return out;
}

/* ---- retained statistics: what a result object keeps of a sample set (R/OptimMCL.R:80-93, R/rsmMCL.R:252-270) ---- */
/* N x d matrix of sufficient statistics (hcar: 4 + 2p columns, else 2 + 2p) with attribute "chains" */
SEXP C_mclb_stats_export(SEXP ctx, SEXP samples, SEXP hcar) {
    mclcar_ctx* c = ctx_of(ctx);
    mclcar_samples* s = samples_of(samples);
    int d = 0;
    CHECK(c, Rf_asLogical(hcar) ? mclcar_hcar_stats(c, s, NULL) : mclcar_dcar_stats(c, s, NULL));
    CHECK(c, mclcar_samples_stats_export(c, s, &d, NULL));
    int64_t N = mclcar_samples_count(s), chains = 0;
    SEXP q = PROTECT(Rf_allocMatrix(REALSXP, (int)N, d));
    int st = mclcar_samples_stats_export(c, s, &d, REAL(q));
    if (st == MCLCAR_OK) st = mclcar_samples_schedule(c, s, &chains, NULL, NULL, NULL);
    SEXP ch = PROTECT(Rf_ScalarReal((double)chains));
    Rf_setAttrib(q, Rf_install("chains"), ch);
    UNPROTECT(2);
    if (st != MCLCAR_OK) Rf_error("%s", mclcar_last_error(c));
    return q;
}
/* a statistics-only sample set from such a matrix, importance point attached */
SEXP C_mclb_samples_from_stats(SEXP ctx, SEXP q, SEXP sites, SEXP chains, SEXP psi, SEXP hcar) {
    mclcar_ctx* c = ctx_of(ctx);
    mclcar_samples* s = NULL;
    CHECK(c, mclcar_samples_from_stats(c, REAL(q), Rf_nrows(q), Rf_ncols(q), (int64_t)Rf_asReal(sites),
                                       (int64_t)Rf_asReal(chains), &s));
    int st = Rf_asLogical(hcar) ? mclcar_hcar_attach(c, s, REAL(psi), LENGTH(psi), NULL)
                                : mclcar_dcar_attach(c, s, REAL(psi), LENGTH(psi), NULL);
    if (st != MCLCAR_OK) { mclcar_samples_free(s); Rf_error("%s", mclcar_last_error(c)); }
    return wrap_samples(s);
}

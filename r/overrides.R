## Drop-in overrides of the hot-path functions of mclcar (R/mcl.dCAR.R, R/mcl.glm.R, R/postZ.R,
## R/rsmSub.R, R/HCAR.sim.R, R/mcl.HCAR.R) behind UNCHANGED signatures: source() after
## library(mclcar), or put in the package's R/ directory together with useDynLib(mclcar).
## NOT RUN IN THIS REPOSITORY (no R in the image).  What is checked here: every .Call below names an
## entry of r/mclcar_b200_glue.c with the right number of arguments, the glue type-checks against
## include/mclcar_b200.h (tests/test_abi_host.py), and the same call sequences are exercised
## through ctypes (mclcar_b200/api.py) by the parity tests.
##
## `data` keeps its reference shape.  Engine contexts (graph + design + observations on the device)
## are cached in the package-level environment .mclb, keyed by a fingerprint of the data, so that
## the outer drivers (OptimMCL, rsmMCL, OptimMCL.HCAR -- R/OptimMCL.R, R/rsmMCL.R) run unmodified.

.mclb <- new.env()
.mclb$control <- list(n.chains = 0, burn = 0L, thin = 0L, keep = TRUE, dtype = 1L, omega = 0)  # dtype 1 = fp32 samples
.mclb$COL_IS_SAMPLE <- 0L   # MCLCAR_COL_IS_SAMPLE: simY (n x N), u.y (K x N)
.mclb$ROW_IS_SAMPLE <- 1L   # MCLCAR_ROW_IS_SAMPLE: Zy (N x n)
.mclb$cache <- list()       # most recently used first; at most .mclb$cache.max contexts stay alive
.mclb$cache.max <- 8L
.mclb$eigen.max <- 2000L    # general graphs up to this size get eigen(W) from R (exact); larger ones the device quadrature

## W -> 0-based CSR slots (p, j, x) WITHOUT densifying: spam (1-based CSR already), Matrix, or a base matrix
.mclb_csr <- function(A) {
    if (inherits(A, "spam"))
        return(list(p = as.integer(A@rowpointers - 1L), j = as.integer(A@colindices - 1L), x = as.numeric(A@entries)))
    if (!inherits(A, "Matrix")) A <- Matrix::Matrix(A, sparse = TRUE)
    R <- methods::as(methods::as(A, "generalMatrix"), "RsparseMatrix")
    list(p = R@p, j = R@j, x = as.numeric(R@x))
}
## cached context whose stored inputs are identical() to `parts` (content check, not a checksum: two data sets with
## equal sums, a permuted y or another n.trial must never share a context); least recently used contexts are dropped
.mclb_cached <- function(kind, parts, make) {
    for (i in seq_along(.mclb$cache)) {
        e <- .mclb$cache[[i]]
        if (identical(e$kind, kind) && identical(e$parts, parts)) {
            .mclb$cache <- c(list(e), .mclb$cache[-i])
            return(e$ctx)
        }
    }
    ctx <- make()
    .mclb$cache <- utils::head(c(list(list(kind = kind, parts = parts, ctx = ctx)), .mclb$cache), .mclb$cache.max)
    ctx          # dropped contexts are released by their finalizers once no sample set refers to them
}
## upload W; the library recognises an exact rook torus in the CSR (tiled sampler, analytic spectrum: no eigen(W)).
## Otherwise: the caller's data$lambda, or eigen() in R up to .mclb$eigen.max sites, or nothing -- the library then
## estimates the spectral sums on the device the first time they are needed (mclcar_graph_estimate_spectrum).
.mclb_graph <- function(ctx, W, lambda) {
    Wt <- .mclb_csr(W)
    .Call(C_mclb_graph_csr, ctx, Wt$p, Wt$j, Wt$x, if (is.null(lambda)) NULL else as.numeric(lambda))
    if (is.null(lambda) && .Call(C_mclb_graph_kind, ctx)[1] == 2L && nrow(W) <= .mclb$eigen.max)
        .Call(C_mclb_set_eigenvalues, ctx, as.numeric(eigen(as.matrix(W), symmetric = TRUE, only.values = TRUE)$values))
    invisible(ctx)
}

## context for the direct CAR model (data$data.vec, data$W) or the GLM (data$y, data$covX, data$W, data$n.trial)
.mclb_ctx <- function(data, glm = FALSE) {
    y <- if (glm) data$y else data$data.vec$y
    X <- if (glm) data$covX else stats::model.matrix(y ~ ., data = data$data.vec)
    nt <- if (glm && !is.null(data$n.trial)) rep_len(as.numeric(data$n.trial), length(y)) else NULL
    .mclb_cached(if (glm) "glm" else "lm", list(y = y, X = X, W = data$W, n.trial = nt, lambda = data$lambda), function() {
        ctx <- .Call(C_mclb_ctx, 0L)
        .mclb_graph(ctx, data$W, data$lambda)
        .Call(C_mclb_model, ctx, X, as.numeric(y), nt)
        ctx
    })
}

## CAR.simTorus(n1, n2, rho, prec) (R/CAR.simLM.R:2-22): W stays SPARSE (the reference returns as.matrix(W), n^2 doubles
## -- 8 TB at 1000 x 1000) and the draw comes from the tiled red-black sampler.  Everything downstream takes the sparse W:
## the overrides upload it as CSR, where the library recognises the torus again.
CAR.simTorus <- function(n1, n2, rho, prec) {
    Wx <- spam::circulant.spam(list(c(2, n2), rep(1, 2)), n = n2)
    W <- spam::kronecker.spam(spam::diag.spam(1, n1), Wx) + spam::kronecker.spam(spam::circulant.spam(list(c(2, n1), rep(1, 2)), n = n1), spam::diag.spam(1, n2))
    list(W = W, X = .Call(C_mclb_sim_torus, as.integer(n1), as.integer(n2), as.numeric(rho), as.numeric(prec)))
}
## CAR.simLM(pars, data) (R/CAR.simLM.R:37-50): one draw of the linear model with CAR errors through the cached context
CAR.simLM <- function(pars, data) {
    sd <- mcl.prep.dCAR(pars, 1L, data)
    as.numeric(.Call(C_mclb_fetch, attr(sd, "ctx"), sd$simY, length(data$data.vec$y), .mclb$COL_IS_SAMPLE))
}

## ---------------------------------------------------------------- direct CAR (R/mcl.dCAR.R)
mcl.prep.dCAR <- function(psi, n.samples, data) {                       # R/mcl.dCAR.R:2-8
    ctx <- .mclb_ctx(data)
    h <- .Call(C_mclb_prep_dcar, ctx, as.numeric(psi), n.samples, .mclb$control)
    structure(list(simY = h, t10 = NA_real_, t20 = NULL), class = "mclb_simdata", ctx = ctx, psi = psi,
              n.samples = n.samples, n = length(data$data.vec$y))
}
## the reference's own list(simY, t10, t20) (e.g. produced on the CPU): upload once; psi0 must be known
.mclb_simdata <- function(data, simdata) {
    if (inherits(simdata, "mclb_simdata")) return(simdata)
    psi <- attr(simdata, "psi")
    if (is.null(psi)) stop("attach the importance point to a CPU-made simdata first: attr(simdata, 'psi') <- psi")
    ctx <- .mclb_ctx(data)
    structure(list(simY = .Call(C_mclb_samples_dcar, ctx, simdata$simY, as.numeric(psi), simdata$t20)),
              class = "mclb_simdata", ctx = ctx, psi = psi, n.samples = ncol(simdata$simY), n = nrow(simdata$simY))
}
mcl.dCAR <- function(pars, data, simdata, rho.cons = c(-0.249, 0.249), Evar = FALSE) {   # R/mcl.dCAR.R:11-50
    simdata <- .mclb_simdata(data, simdata)
    r <- .Call(C_mclb_mcl_dcar, attr(simdata, "ctx"), simdata$simY, matrix(as.numeric(pars), nrow = 1), rho.cons, NULL, NULL)
    if (Evar) as.numeric(r) else r[1, 1]
}
mcl.profile.dCAR <- function(rho, data, simdata, rho.cons = c(-0.249, 0.249), Evar = FALSE)   # R/mcl.dCAR.R:53-96
    mcl.dCAR(c(rho, sigmabeta(rho, data)), data, simdata, rho.cons, Evar)
mc.gradlik.dCAR <- function(pars, data, simdata, Evar = 0) {            # R/mcl.dCAR.R:213-283
    simdata <- .mclb_simdata(data, simdata)
    r <- .Call(C_mclb_grad_dcar, attr(simdata, "ctx"), simdata$simY, as.numeric(pars), as.integer(Evar))
    if (Evar == 0) r[[1]] else list(grad = r[[1]], grad.var = r[[2]])
}
mc.Hesslik.dCAR <- function(pars, data, simdata) {                      # R/mcl.dCAR.R:287-394
    simdata <- .mclb_simdata(data, simdata)
    .Call(C_mclb_hess_dcar, attr(simdata, "ctx"), simdata$simY, as.numeric(pars), 0L)
}
MCMLE.var <- function(pars, data, simdata) {                            # R/mcl.dCAR.R:156-165
    simdata <- .mclb_simdata(data, simdata)
    .Call(C_mclb_hess_dcar, attr(simdata, "ctx"), simdata$simY, as.numeric(pars), 1L)
}

## design points and path points in ONE device call each instead of a loop of mcl.dCAR calls
Exp.dCAR <- function(exp.design, data, psi, n0, mc.data) {              # R/rsmSub.R:65-97
    if (is.null(mc.data)) stop("Please provide the Monte Carlo samples by specifying mc.data!")
    mc.data <- .mclb_simdata(data, mc.data)
    n.samples <- attr(mc.data, "n.samples"); nb.sim <- floor(n.samples / n0)
    dd <- rsm::decode.data(exp.design); K <- nrow(exp.design)
    pars <- cbind(dd$rho, dd$sigma, matrix(as.numeric(psi[-c(1, 2)]), K, length(psi) - 2, byrow = TRUE))
    subs <- lapply(seq_len(K), function(k) if (exp.design$std.order[k] > 4) sample(n.samples, nb.sim) - 1 else numeric(0))
    ptr <- c(0, cumsum(lengths(subs)))
    r <- .Call(C_mclb_mcl_dcar, attr(mc.data, "ctx"), mc.data$simY, pars, NULL, as.numeric(unlist(subs)), as.numeric(ptr))
    exp.design$mc.lr <- r[, 1]; exp.design$mc.var <- r[, 2]
    exp.design
}
ExpPath.dCAR <- function(exp.design, data, psi, mc.data) {              # R/rsmSub.R:102-115
    if (is.null(mc.data)) stop("Please provide the Monte Carlo samples by specifying mc.data!")
    mc.data <- .mclb_simdata(data, mc.data)
    K <- nrow(exp.design)
    pars <- cbind(exp.design$rho, exp.design$sigma, matrix(as.numeric(psi[-c(1, 2)]), K, length(psi) - 2, byrow = TRUE))
    ## the reference's `rho.cons = NULL` sits inside c(...) (:110) and is lost: the default guard applies
    r <- .Call(C_mclb_mcl_dcar, attr(mc.data, "ctx"), mc.data$simY, pars, c(-0.249, 0.249), NULL, NULL)
    exp.design$mc.lr <- r[, 1]; exp.design$mc.var <- r[, 2]
    exp.design
}

## ---------------------------------------------------------------- GLM with latent CAR (R/mcl.glm.R, R/postZ.R)
mcl.prep.glm <- function(data, family, psi, Z.start = NULL, mcmc.control = list(), pilot.run = TRUE,
                         pilot.plot = FALSE, plot.diag = FALSE) {       # R/mcl.glm.R:5-40
    ## Z.start, pilot.run and the MALA scale have no role: parallel Metropolis-within-Gibbs chains, no scale to tune
    con <- list(method = "mala", N.Zy = 1e4, thin = 5, burns = 5e3); con[names(mcmc.control)] <- mcmc.control
    ctx <- .mclb_ctx(data, glm = TRUE)
    fam <- c(binom = 1L, poisson = 2L)[[family]]
    n.samples <- (con$N.Zy - con$burns) %/% con$thin                    # rows of Z, R/postZsub.R:427
    data$Zy <- .Call(C_mclb_prep_glm, ctx, fam, as.numeric(psi), n.samples, .mclb$control)
    data$lHZy.psi <- NULL; data$mcmc.pars <- con
    structure(data, class = "mclb_mcdata", ctx = ctx, n.samples = n.samples)
}
postZ <- function(data, family, psi, Z.start, mcmc.control = list(), plots = FALSE) {    # R/postZ.R:2-14
    md <- mcl.prep.glm(data, family, psi, Z.start, mcmc.control, pilot.run = FALSE)
    ## Z: N x n, row = sample (R/postZsub.R:427); AC.r: mean acceptance rate of the chains
    list(Z = .Call(C_mclb_fetch, attr(md, "ctx"), md$Zy, length(data$y), .mclb$ROW_IS_SAMPLE),
         AC.r = .Call(C_mclb_samples_diag, attr(md, "ctx"), md$Zy)[1])
}
mcl.glm <- function(pars, mcdata, family, Evar = FALSE) {               # R/mcl.glm.R:43-47
    r <- .Call(C_mclb_mcl_glm, attr(mcdata, "ctx"), mcdata$Zy, matrix(as.numeric(pars), nrow = 1), NULL, NULL)
    if (Evar) as.numeric(r) else r[1, 1]
}
mcl.grad.glm <- function(pars, mcdata, family, Evar = FALSE) {          # R/mcl.glm.R:78-82
    r <- .Call(C_mclb_grad_glm, attr(mcdata, "ctx"), mcdata$Zy, as.numeric(pars), Evar)
    if (Evar) list(grad.lr = r[[1]], grad.var = r[[2]]) else r[[1]]
}
mcl.Hessian.glm <- function(pars, mcdata, family)                       # R/mcl.glm.R:85-89
    .Call(C_mclb_hess_glm, attr(mcdata, "ctx"), mcdata$Zy, as.numeric(pars))
get.beta.glm <- function(beta0, rho.sig, mcdata, family) {              # R/mcl.glm.R:50-54, R/mcl.bin.R:83-95
    r <- .Call(C_mclb_get_beta_glm, attr(mcdata, "ctx"), mcdata$Zy, as.numeric(rho.sig), as.numeric(beta0), 2L, 1)
    list(x = r[[1]], fvec = r[[2]], termcd = r[[3]], iter = r[[4]])
}
mcl.profile.glm <- function(pars, beta0, mcdata, family, Evar = FALSE) {   # R/mcl.glm.R:57-61, R/mcl.bin.R:99-156
    r <- .Call(C_mclb_get_beta_glm, attr(mcdata, "ctx"), mcdata$Zy, as.numeric(pars[1:2]), as.numeric(beta0), 3L, 1)
    lr <- mcl.glm(c(pars[1:2], r[[1]]), mcdata, family, Evar = TRUE)
    c(if (Evar) lr else lr[1], r[[1]], r[[2]])
}
Exp.CAR.glm <- function(exp.design, data, family, psi, n0, mc.data) {   # R/rsmSub.R:159-192
    if (is.null(mc.data)) stop("Please provide the Monte Carlo samples by specifying mc.data!")
    n.samples <- attr(mc.data, "n.samples"); nb.sim <- floor(n.samples / n0)
    dd <- rsm::decode.data(exp.design); K <- nrow(exp.design)
    pars <- cbind(dd$rho, dd$sigma, matrix(as.numeric(psi[-c(1, 2)]), K, length(psi) - 2, byrow = TRUE))
    subs <- lapply(seq_len(K), function(k) if (exp.design$std.order[k] > 4) sample(n.samples, nb.sim) - 1 else numeric(0))
    r <- .Call(C_mclb_mcl_glm, attr(mc.data, "ctx"), mc.data$Zy, pars, as.numeric(unlist(subs)), as.numeric(c(0, cumsum(lengths(subs)))))
    exp.design$mc.lr <- r[, 1]
    exp.design$mc.var <- ifelse(r[, 2] == 0, 1e-8, pmin(1e8, r[, 2]))    # per row, as R/rsmSub.R:183,188
    exp.design
}
ExpPath.CAR.glm <- function(exp.design, data, family, psi, mc.data) {   # R/rsmSub.R:197-212
    if (is.null(mc.data)) stop("Please provide the Monte Carlo samples by specifying mc.data!")
    K <- nrow(exp.design)
    pars <- cbind(exp.design$rho, exp.design$sigma, matrix(as.numeric(psi[-c(1, 2)]), K, length(psi) - 2, byrow = TRUE))
    r <- .Call(C_mclb_mcl_glm, attr(mc.data, "ctx"), mc.data$Zy, pars, NULL, NULL)
    exp.design$mc.lr <- r[, 1]
    exp.design$mc.var <- ifelse(r[1, 2] == 0, 1e-8, min(1e8, r[1, 2]))   # single element, as coded at R/rsmSub.R:210
    exp.design
}

## ---------------------------------------------------------------- hierarchical CAR (R/HCAR.sim.R, R/mcl.HCAR.R)
## data: list or environment with y, X, W (or NULL), M, Z (and optionally aa = eigen(W), bb = eigen(M), R/mcl.HCAR.R:166-173)
.mclb_hcar_ctx <- function(data) {
    g <- function(nm) if (is.environment(data)) get0(nm, envir = data) else data[[nm]]
    y <- g("y"); X <- g("X"); W <- g("W"); M <- g("M"); Z <- g("Z")
    .mclb_cached("hcar", list(y = y, X = X, W = W, M = M, Z = Z, aa = g("aa"), bb = g("bb")), function() {
        ctx <- .Call(C_mclb_ctx, 0L)
        if (is.null(W)) .Call(C_mclb_graph_csr, ctx, integer(length(y) + 1), integer(0), numeric(0), NULL)    # the W = NULL variant
        else .mclb_graph(ctx, W, g("aa"))
        .Call(C_mclb_model, ctx, as.matrix(X), as.numeric(y), NULL)
        Mt <- .mclb_csr(M); Zt <- .mclb_csr(Z)
        bb <- if (is.null(g("bb"))) eigen(as.matrix(M), symmetric = TRUE, only.values = TRUE)$values else g("bb")   # K x K: districts
        .Call(C_mclb_hcar_model, ctx, Mt$p, Mt$j, Mt$x, Zt$p, Zt$j, Zt$x, as.numeric(bb), !is.null(W))
        ctx
    })
}
sim.HCAR <- function(psi, data, n.samples) {                            # R/HCAR.sim.R:3-87
    ctx <- .mclb_hcar_ctx(data)
    h <- .Call(C_mclb_sim_hcar, ctx, as.numeric(psi), n.samples, .mclb$control)
    structure(list(psi = psi, n.samples = n.samples, u.y = h), class = "mclb_hcar", ctx = ctx)
}
mcl.HCAR <- function(pars, mcdata, data, Evar = FALSE) {                # R/mcl.HCAR.R:4-97
    r <- .Call(C_mclb_mcl_hcar, attr(mcdata, "ctx"), mcdata$u.y, matrix(as.numeric(pars), nrow = 1))
    if (Evar) as.numeric(r) else r[1, 1]
}
mcl.grad.HCAR <- function(pars, mcdata, data, Evar = FALSE) {           # R/mcl.HCAR.R:101-213
    r <- .Call(C_mclb_grad_hcar, attr(mcdata, "ctx"), mcdata$u.y, as.numeric(pars), Evar)
    if (Evar) list(grad.lr = r[[1]], grad.var = r[[2]]) else r[[1]]
}
mcl.Hessian.HCAR <- function(pars, mcdata, data)                        # R/mcl.HCAR.R:216-391
    .Call(C_mclb_hess_hcar, attr(mcdata, "ctx"), mcdata$u.y, as.numeric(pars))


## ---------------------------------------------------------------- profile / exact pieces and summaries (SURVEY.md 8f)
sigmabeta <- function(rho, data) {                                      # R/loglik.dCAR.R:72-86
    ctx <- .mclb_ctx(data)
    as.numeric(.Call(C_mclb_sigmabeta, ctx, as.numeric(rho), ncol(stats::model.matrix(y ~ ., data = data$data.vec)))[1, ])
}
loglik.dCAR <- function(pars, data, rho.cons = c(-0.249, 0.249))        # R/loglik.dCAR.R:7-42
    .Call(C_mclb_loglik_dcar, .mclb_ctx(data), matrix(as.numeric(pars), nrow = 1), rho.cons)
Hessian.dCAR <- function(pars, data)                                    # R/loglik.dCAR.R:171-219
    .Call(C_mclb_hessian_dcar, .mclb_ctx(data), as.numeric(pars))
## the exact backdrop of plot.rsmMCL -- R/rsmMCL.R:95-102 apply()s loglik.dCAR over the 100 x 100 grid -- in ONE call;
## same list shape as Exacts[[i]]
exact.surface.dCAR <- function(r.vec, s.vec, psi, data) {
    g <- as.matrix(expand.grid(r.vec, s.vec))
    pars <- cbind(g, matrix(as.numeric(psi[-c(1, 2)]), nrow(g), length(psi) - 2, byrow = TRUE))
    ctx <- .mclb_ctx(data)
    lrs <- .Call(C_mclb_loglik_dcar, ctx, pars, c(-0.249, 0.249)) - .Call(C_mclb_loglik_dcar, ctx, matrix(as.numeric(psi), nrow = 1), c(-0.249, 0.249))
    list(r.vec = r.vec, s.vec = s.vec, lrs = matrix(lrs, length(r.vec), length(s.vec)))
}
## the MONTE CARLO likelihood surface on the same grid: what the rsm design points sample from.  Returned in the shape
## of Exacts[[i]] (plus $vars), so plot.RSM(..., exact = mc.surface(...), exact.vals = TRUE) draws it as the image
## behind the fitted response surface -- for the GLM families this is the only backdrop there is (R/rsmMCL.R:106).
mc.surface <- function(r.vec, s.vec, psi, mc.data, data = NULL, family = "gauss", rho.cons = NULL) {
    if (family == "gauss") mc.data <- .mclb_simdata(data, mc.data)
    h <- if (family == "gauss") mc.data$simY else mc.data$Zy
    fam <- c(gauss = 0L, binom = 1L, poisson = 2L)[[family]]
    beta <- if (length(psi) > 2) as.numeric(psi[-c(1, 2)]) else NULL
    r <- .Call(C_mclb_surface, attr(mc.data, "ctx"), h, fam, as.numeric(r.vec), as.numeric(s.vec), beta, rho.cons)
    list(r.vec = r.vec, s.vec = s.vec, lrs = r[[1]], vars = r[[2]])
}
## plot.rsmMCL with the Monte Carlo surface of each iteration's sample set as the backdrop (mc.surface = TRUE), on the
## grid the exact values use; everything else is the reference's plot (R/plot.rsmMCL.R:2-88)
plot.rsmMCL <- function(x, family, trace.all = TRUE, plain = TRUE, mc.surface = FALSE, length = 100, ...) {
    if (mc.surface) {
        Psis <- rbind(x$psi0, do.call(rbind, x$Psi))
        for (i in seq_len(if (trace.all) x$N.iter else 1L)) {
            sim <- x$Sim.data[[i]]
            if (inherits(sim, "mclb_retained")) sim <- restore.mcdata(sim, x$data)
            r.vec <- if (is.list(x$Exacts[[i]])) x$Exacts[[i]]$r.vec else seq(-0.24, 0.24, length.out = length)
            s.vec <- if (is.list(x$Exacts[[i]])) x$Exacts[[i]]$s.vec else seq(0.5, 2, length.out = length) * Psis[i, 2]
            x$Exacts[[i]] <- mc.surface(r.vec, s.vec, Psis[i, ], sim, x$data, family)
        }
        family <- "gauss"          # plot.exact <- family == "gauss" (R/plot.rsmMCL.R:3): draw the backdrop
    }
    get("plot.rsmMCL", envir = asNamespace("mclcar"))(x, family, trace.all, plain, ...)
}

## ---------------------------------------------------------------- retained statistics instead of retained matrices
## OptimMCL / rsmMCL keep every iteration's sample set in MC.datas / Sim.data (R/OptimMCL.R:80-93, R/rsmMCL.R:252-270)
## for summary() to re-evaluate vmle.* on (R/summary.OptimMCL.R:24-26).  A device handle does not survive saveRDS();
## retain.mcdata() turns one into a plain R object holding the N x d sufficient statistics (N * d doubles instead of
## N * n), restore.mcdata() rebuilds a statistics-only sample set from it -- every mcl.* / vmle.* function works on it.
retain.mcdata <- function(mc.data) {
    hcar <- inherits(mc.data, "mclb_hcar")
    h <- if (hcar) mc.data$u.y else mc.data$simY
    q <- .Call(C_mclb_stats_export, attr(mc.data, "ctx"), h, hcar)
    structure(list(q = q, psi = if (hcar) mc.data$psi else attr(mc.data, "psi"), hcar = hcar,
                   n = attr(mc.data, "n"), n.samples = nrow(q)), class = "mclb_retained")
}
restore.mcdata <- function(kept, data) {
    ctx <- if (kept$hcar) .mclb_hcar_ctx(data) else .mclb_ctx(data)
    sites <- if (kept$hcar) { M <- if (is.environment(data)) get("M", envir = data) else data$M; nrow(M) } else kept$n
    h <- .Call(C_mclb_samples_from_stats, ctx, kept$q, sites, attr(kept$q, "chains"), as.numeric(kept$psi), kept$hcar)
    if (kept$hcar) structure(list(psi = kept$psi, n.samples = kept$n.samples, u.y = h), class = "mclb_hcar", ctx = ctx)
    else structure(list(simY = h, t10 = NA_real_, t20 = NULL), class = "mclb_simdata", ctx = ctx, psi = kept$psi,
                   n.samples = kept$n.samples, n = kept$n)
}

## ---- Poisson-response hierarchical model (an EXTENSION: the reference's sim.HCAR / mcl.HCAR are Gaussian, R/HCAR.sim.R:42-45)
##   y_i | u ~ Poisson(exp(x_i' beta + u_{d(i)})),  u ~ N(0, sigma (I - rho M)^-1);  psi = (rho, sigma, beta)
## Summed over the individuals of a district this is the Poisson GLM of R/mcl.pois.R on the K districts with
## observations Y_k = sum y_i and the offset o_k(beta) = log sum exp(x_i' beta) in the place of X beta, up to the
## constant sum_i y_i x_i' beta - sum_k Y_k o_k(beta): the same .Call entries as the GLM path, nothing new on the device.
## data: list(y, X, district (1..K), M, bb = eigen(M)$values)
.mclb_pois_offset <- function(beta, data) {
    eta <- as.numeric(data$X %*% beta)
    top <- tapply(eta, data$district, max)
    as.numeric(top + log(tapply(exp(eta - top[data$district]), data$district, sum)))
}
.mclb_pois_const <- function(beta, data, off) sum(data$y * (data$X %*% beta)) - sum(tapply(data$y, data$district, sum) * off)
sim.HCAR.pois <- function(psi, data, n.samples) {
    Y <- as.numeric(tapply(data$y, data$district, sum))
    o0 <- .mclb_pois_offset(psi[-(1:2)], data)
    ctx <- .Call(C_mclb_ctx, 0L)                 # its own context: the design is replaced for every batch of candidates
    .mclb_graph(ctx, data$M, data$bb)
    .Call(C_mclb_model, ctx, matrix(o0, ncol = 1), Y, NULL)
    h <- .Call(C_mclb_prep_glm, ctx, 2L, c(psi[1:2], 1), n.samples, .mclb$control)
    structure(list(u.y = h, psi = psi, Y = Y), ctx = ctx)
}
mcl.HCAR.pois <- function(pars, mcdata, data) {                             # pars: K x (2 + p) -> K x 2 (lr, v.lr)
    pars <- matrix(as.numeric(pars), ncol = length(mcdata$psi))
    ctx <- attr(mcdata, "ctx"); beta0 <- mcdata$psi[-(1:2)]
    o0 <- .mclb_pois_offset(beta0, data); c0 <- .mclb_pois_const(beta0, data, o0)
    key <- apply(pars[, -(1:2), drop = FALSE], 1, paste, collapse = " ")
    same0 <- key == paste(beta0, collapse = " ")
    out <- matrix(NA_real_, nrow(pars), 2)
    todo <- unique(key[!same0]); groups <- if (length(todo)) split(todo, ceiling(seq_along(todo) / 7)) else list(character(0))
    first <- TRUE
    for (g in groups) {
        rows.g <- lapply(g, function(k) which(key == k))
        offs <- lapply(rows.g, function(r) .mclb_pois_offset(pars[r[1], -(1:2)], data))
        Xo <- do.call(cbind, c(list(o0), offs))
        .Call(C_mclb_model, ctx, Xo, mcdata$Y, NULL)
        .Call(C_mclb_attach_glm, ctx, mcdata$u.y, 2L, c(mcdata$psi[1:2], 1, rep(0, length(g))))
        rows <- c(if (first) which(same0), unlist(rows.g)); if (!length(rows)) next
        cand <- matrix(0, length(rows), 2 + ncol(Xo)); cand[, 1:2] <- pars[rows, 1:2]; shift <- numeric(length(rows))
        for (t in seq_along(rows)) {
            j <- match(key[rows[t]], g)
            if (is.na(j)) cand[t, 3] <- 1 else {
                cand[t, 3 + j] <- 1
                shift[t] <- .mclb_pois_const(pars[rows[t], -(1:2)], data, offs[[j]]) - c0
            }
        }
        r <- .Call(C_mclb_mcl_glm, ctx, mcdata$u.y, cand, NULL, NULL)
        out[rows, ] <- cbind(r[, 1] + shift, r[, 2]); first <- FALSE
    }
    .Call(C_mclb_model, ctx, matrix(o0, ncol = 1), mcdata$Y, NULL)
    .Call(C_mclb_attach_glm, ctx, mcdata$u.y, 2L, c(mcdata$psi[1:2], 1))
    out
}
mcl.grad.HCAR.pois <- function(pars, mcdata, data) {                        # gradient by the chain rule through the district offsets
    ctx <- attr(mcdata, "ctx"); p <- length(pars) - 2L; beta <- pars[-(1:2)]
    o0 <- .mclb_pois_offset(mcdata$psi[-(1:2)], data); o1 <- .mclb_pois_offset(beta, data)
    w <- exp(as.numeric(data$X %*% beta) - o1[data$district])              # e^{x_i' beta} / S_k
    xbar <- sapply(seq_len(p), function(l) as.numeric(tapply(w * data$X[, l], data$district, sum)))
    .Call(C_mclb_model, ctx, cbind(o0, o1, xbar), mcdata$Y, NULL)
    .Call(C_mclb_attach_glm, ctx, mcdata$u.y, 2L, c(mcdata$psi[1:2], 1, 0, rep(0, p)))
    g <- .Call(C_mclb_grad_glm, ctx, mcdata$u.y, c(pars[1:2], 0, 1, rep(0, p)), FALSE)[[1]]
    .Call(C_mclb_model, ctx, matrix(o0, ncol = 1), mcdata$Y, NULL)
    .Call(C_mclb_attach_glm, ctx, mcdata$u.y, 2L, c(mcdata$psi[1:2], 1))
    c(g[1:2], g[-(1:4)] + as.numeric(crossprod(data$X, data$y)) - as.numeric(crossprod(xbar, mcdata$Y)))
}

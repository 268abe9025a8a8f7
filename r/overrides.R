## Drop-in overrides of the hot-path functions of mclcar (R/mcl.dCAR.R, R/mcl.glm.R, R/HCAR.sim.R,
## R/mcl.HCAR.R) behind UNCHANGED signatures: source() after library(mclcar), or put in the
## package's R/ directory together with useDynLib(mclcar).  NOT RUN IN THIS REPOSITORY (no R in
## the image); the same call sequences are exercised through ctypes in tests/.
##
## `data` keeps its reference shape; the engine context is cached in an attribute so that the
## outer drivers (OptimMCL, rsmMCL -- R/OptimMCL.R, R/rsmMCL.R) run unmodified.

.mclb <- new.env()
.mclb$control <- list(n.chains = 0, burn = 0L, thin = 0L, keep = TRUE, dtype = 1L, omega = 0)
.mclb$COL_IS_SAMPLE <- 0L   # MCLCAR_COL_IS_SAMPLE: simY (n x N), u.y (K x N)
.mclb$ROW_IS_SAMPLE <- 1L   # MCLCAR_ROW_IS_SAMPLE: Zy (N x n)

.mclb_ctx <- function(data, glm = FALSE) {
    key <- if (glm) "glm" else "lm"
    if (!is.null(attr(data, "mclb"))) return(attr(data, "mclb"))
    ctx <- .Call(C_mclb_ctx, 0L)
    W <- data$W
    Wt <- methods::as(Matrix::Matrix(W, sparse = TRUE), "RsparseMatrix")   # CSR, 0-based
    lambda <- if (is.null(data$lambda)) eigen(as.matrix(W), symmetric = TRUE, only.values = TRUE)$values else data$lambda
    .Call(C_mclb_graph_csr, ctx, Wt@p, Wt@j, Wt@x, lambda)
    if (glm) .Call(C_mclb_model, ctx, data$covX, as.numeric(data$y), rep_len(as.numeric(data$n.trial), length(data$y)))
    else .Call(C_mclb_model, ctx, model.matrix(y ~ ., data = data$data.vec), data$data.vec$y, NULL)
    ctx
}

mcl.prep.dCAR <- function(psi, n.samples, data) {                       # R/mcl.dCAR.R:2-8
    ctx <- .mclb_ctx(data)
    h <- .Call(C_mclb_prep_dcar, ctx, as.numeric(psi), n.samples, .mclb$control)
    structure(list(simY = h, t10 = NA_real_, t20 = NULL), class = "mclb_simdata", ctx = ctx, n = length(data$data.vec$y))
}
mcl.dCAR <- function(pars, data, simdata, rho.cons = c(-0.249, 0.249), Evar = FALSE) {   # R/mcl.dCAR.R:11-50
    if (!inherits(simdata, "mclb_simdata")) {   # the reference's own list(simY, t10, t20): upload once
        ctx <- .mclb_ctx(data)
        simdata <- structure(list(simY = .Call(C_mclb_samples_dcar, ctx, simdata$simY, attr(simdata, "psi"), simdata$t20)),
                             class = "mclb_simdata", ctx = ctx)
    }
    r <- .Call(C_mclb_mcl_dcar, attr(simdata, "ctx"), simdata$simY, matrix(as.numeric(pars), nrow = 1), rho.cons, NULL, NULL)
    if (Evar) as.numeric(r) else r[1, 1]
}
mcl.profile.dCAR <- function(rho, data, simdata, rho.cons = c(-0.249, 0.249), Evar = FALSE)   # R/mcl.dCAR.R:53-96
    mcl.dCAR(c(rho, sigmabeta(rho, data)), data, simdata, rho.cons, Evar)
mc.gradlik.dCAR <- function(pars, data, simdata, Evar = 0) {            # R/mcl.dCAR.R:213-283
    r <- .Call(C_mclb_grad_dcar, attr(simdata, "ctx"), simdata$simY, as.numeric(pars), as.integer(Evar))
    if (Evar == 0) r[[1]] else list(grad = r[[1]], grad.var = r[[2]])
}
mc.Hesslik.dCAR <- function(pars, data, simdata)                        # R/mcl.dCAR.R:287-394
    .Call(C_mclb_hess_dcar, attr(simdata, "ctx"), simdata$simY, as.numeric(pars), 0L)
MCMLE.var <- function(pars, data, simdata)                              # R/mcl.dCAR.R:156-165
    .Call(C_mclb_hess_dcar, attr(simdata, "ctx"), simdata$simY, as.numeric(pars), 1L)

## batched design-point evaluation replacing the loops of Exp.dCAR / ExpPath.dCAR (R/rsmSub.R:78-95,110)
Exp.dCAR <- function(exp.design, data, psi, n0, mc.data) {
    if (is.null(mc.data)) stop("Please provide the Monte Carlo samples by specifying mc.data!")
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
    K <- nrow(exp.design)
    pars <- cbind(exp.design$rho, exp.design$sigma, matrix(as.numeric(psi[-c(1, 2)]), K, length(psi) - 2, byrow = TRUE))
    ## the reference's `rho.cons = NULL` sits inside c(...) (:110) and is lost: the default guard applies
    r <- .Call(C_mclb_mcl_dcar, attr(mc.data, "ctx"), mc.data$simY, pars, c(-0.249, 0.249), NULL, NULL)
    exp.design$mc.lr <- r[, 1]; exp.design$mc.var <- r[, 2]
    exp.design
}

mcl.prep.glm <- function(data, family, psi, Z.start = NULL, mcmc.control = list(), pilot.run = TRUE,
                         pilot.plot = FALSE, plot.diag = FALSE) {       # R/mcl.glm.R:5-40
    con <- list(N.Zy = 1e4, thin = 5, burns = 5e3); con[names(mcmc.control)] <- mcmc.control
    ctx <- .mclb_ctx(data, glm = TRUE)
    fam <- c(binom = 1L, poisson = 2L)[[family]]
    data$Zy <- .Call(C_mclb_prep_glm, ctx, fam, as.numeric(psi), (con$N.Zy - con$burns) %/% con$thin, .mclb$control)
    data$lHZy.psi <- NULL; data$mcmc.pars <- con
    structure(data, class = "mclb_mcdata", ctx = ctx)
}
mcl.glm <- function(pars, mcdata, family, Evar = FALSE) {               # R/mcl.glm.R:43-47
    r <- .Call(C_mclb_mcl_glm, attr(mcdata, "ctx"), mcdata$Zy, matrix(as.numeric(pars), nrow = 1), NULL, NULL)
    if (Evar) as.numeric(r) else r[1, 1]
}
postZ <- function(data, family, psi, Z.start, mcmc.control = list(), plots = FALSE) {    # R/postZ.R:2-14
    md <- mcl.prep.glm(data, family, psi, Z.start, mcmc.control, pilot.run = FALSE)
    ## Z: N x n, row = sample (R/postZsub.R:427); AC.r: mean acceptance rate of the Metropolis-within-Gibbs chains
    list(Z = .Call(C_mclb_fetch, attr(md, "ctx"), md$Zy, length(data$y), .mclb$ROW_IS_SAMPLE),
         AC.r = .Call(C_mclb_samples_diag, attr(md, "ctx"), md$Zy)[1])
}
get.beta.glm <- function(beta0, rho.sig, mcdata, family) {              # R/mcl.glm.R:50-54, R/mcl.bin.R:83-95
    r <- .Call(C_mclb_get_beta_glm, attr(mcdata, "ctx"), mcdata$Zy, as.numeric(rho.sig), as.numeric(beta0), 2L, 1)
    list(x = r[[1]], fvec = r[[2]], termcd = r[[3]], iter = r[[4]])
}
mcl.profile.glm <- function(pars, beta0, mcdata, family, Evar = FALSE) {   # R/mcl.glm.R:57-61, R/mcl.bin.R:99-156
    r <- .Call(C_mclb_get_beta_glm, attr(mcdata, "ctx"), mcdata$Zy, as.numeric(pars[1:2]), as.numeric(beta0), 3L, 1)
    lr <- mcl.glm(c(pars[1:2], r[[1]]), mcdata, family, Evar = TRUE)
    c(if (Evar) lr else lr[1], r[[1]], r[[2]])
}
## batched design-point evaluation replacing the loops of Exp.CAR.glm / ExpPath.CAR.glm (R/rsmSub.R:159-212)
Exp.CAR.glm <- function(exp.design, data, family, psi, n0, mc.data) {
    if (is.null(mc.data)) stop("Please provide the Monte Carlo samples by specifying mc.data!")
    n.samples <- (mc.data$mcmc.pars$N.Zy - mc.data$mcmc.pars$burns) %/% mc.data$mcmc.pars$thin
    nb.sim <- floor(n.samples / n0)
    dd <- rsm::decode.data(exp.design); K <- nrow(exp.design)
    pars <- cbind(dd$rho, dd$sigma, matrix(as.numeric(psi[-c(1, 2)]), K, length(psi) - 2, byrow = TRUE))
    subs <- lapply(seq_len(K), function(k) if (exp.design$std.order[k] > 4) sample(n.samples, nb.sim) - 1 else numeric(0))
    r <- .Call(C_mclb_mcl_glm, attr(mc.data, "ctx"), mc.data$Zy, pars, as.numeric(unlist(subs)), as.numeric(c(0, cumsum(lengths(subs)))))
    exp.design$mc.lr <- r[, 1]
    exp.design$mc.var <- ifelse(r[, 2] == 0, 1e-8, pmin(1e8, r[, 2]))    # per row, as R/rsmSub.R:183,188
    exp.design
}
ExpPath.CAR.glm <- function(exp.design, data, family, psi, mc.data) {
    if (is.null(mc.data)) stop("Please provide the Monte Carlo samples by specifying mc.data!")
    K <- nrow(exp.design)
    pars <- cbind(exp.design$rho, exp.design$sigma, matrix(as.numeric(psi[-c(1, 2)]), K, length(psi) - 2, byrow = TRUE))
    r <- .Call(C_mclb_mcl_glm, attr(mc.data, "ctx"), mc.data$Zy, pars, NULL, NULL)
    exp.design$mc.lr <- r[, 1]
    exp.design$mc.var <- ifelse(r[1, 2] == 0, 1e-8, min(1e8, r[1, 2]))   # single element, as coded at R/rsmSub.R:210
    exp.design
}
mcl.grad.glm <- function(pars, mcdata, family, Evar = FALSE) {          # R/mcl.glm.R:78-82
    r <- .Call(C_mclb_grad_glm, attr(mcdata, "ctx"), mcdata$Zy, as.numeric(pars), Evar)
    if (Evar) list(grad.lr = r[[1]], grad.var = r[[2]]) else r[[1]]
}
mcl.Hessian.glm <- function(pars, mcdata, family)                       # R/mcl.glm.R:85-89
    .Call(C_mclb_hess_glm, attr(mcdata, "ctx"), mcdata$Zy, as.numeric(pars))

sim.HCAR <- function(psi, data, n.samples) {                            # R/HCAR.sim.R:3-87
    ctx <- attr(data, "mclb"); stopifnot(!is.null(ctx))                 # built once by mclb.hcar.data(data)
    h <- .Call(C_mclb_sim_hcar, ctx, as.numeric(psi), n.samples, .mclb$control)
    structure(list(psi = psi, n.samples = n.samples, u.y = h), class = "mclb_hcar", ctx = ctx)
}
mcl.HCAR <- function(pars, mcdata, data, Evar = FALSE) {                # R/mcl.HCAR.R:4-97
    r <- .Call(C_mclb_mcl_hcar, attr(mcdata, "ctx"), mcdata$u.y, matrix(as.numeric(pars), nrow = 1))
    if (Evar) as.numeric(r) else r[1, 1]
}

#!/usr/bin/env Rscript
# pin_with_cmdstanr.R -- evaluates the reference's Stan programs with cmdstanr at the vectors of the pin kit.
#
#   Rscript scripts/pin_with_cmdstanr.R <dir holding approxGP.stan, approxGP2.stan, approxGP3.stan, approxGP4.stan> [tests/golden/pin]
#
# (<dir> = system.file(package = "bayesVG") for an installed bayesVG, or the inst/ directory of a checkout.)  For every
# <name>.data.json / <name>.theta.json pair written by tests/golden/make_pin_kit.py this compiles the program named in the
# theta file exactly as R/findSpatiallyVariableFeaturesBayes.R:337-341 does (stanc O1), instantiates it on the data, and
# writes <name>.cmdstanr.json with $log_prob() and $grad_log_prob() at every vector, jacobian = FALSE and TRUE.
# tests/test_pin_cmdstanr.py picks those files up: the CPU suite compares the oracle with them, the GPU suite the CUDA engine
# (bvg_log_prob is the parity hook of include/bayesvg_b200.h).  Needs: R, cmdstanr (>= 0.6 for the model methods), CmdStan.
suppressPackageStartupMessages({
  library(cmdstanr)
  library(jsonlite)
})
args <- commandArgs(trailingOnly = TRUE)
if (length(args) < 1) stop("usage: pin_with_cmdstanr.R <dir with approxGP*.stan> [kit dir]")
stan_dir <- args[1]
kit <- if (length(args) > 1) args[2] else file.path("tests", "golden", "pin")
index <- fromJSON(file.path(kit, "index.json"), simplifyDataFrame = TRUE)
models <- list()
for (i in seq_len(nrow(index))) {
  name <- index$name[i]
  prog <- index$program[i]
  if (is.null(models[[prog]])) {
    mod <- cmdstan_model(file.path(stan_dir, prog), compile = FALSE)
    mod$compile(stanc_options = list("O1"), force_recompile = TRUE, compile_model_methods = TRUE)
    models[[prog]] <- mod
  }
  mod <- models[[prog]]
  data_file <- file.path(kit, paste0(name, ".data.json"))
  thetas <- fromJSON(file.path(kit, paste0(name, ".theta.json")))$thetas
  if (is.null(dim(thetas))) thetas <- matrix(thetas, nrow = 1)
  # any fit object carries the model methods; one fixed-parameter draw is the cheapest
  fit <- mod$sample(data = data_file, chains = 1, iter_warmup = 0, iter_sampling = 1, fixed_param = TRUE,
                    init = 0, refresh = 0, show_messages = FALSE)
  fit$init_model_methods()
  out <- list(name = name, program = prog, cmdstan_version = cmdstan_version(),
              cmdstanr_version = as.character(utils::packageVersion("cmdstanr")),
              lp_j0 = numeric(0), lp_j1 = numeric(0), grad_j0 = list(), grad_j1 = list())
  for (r in seq_len(nrow(thetas))) {
    th <- as.numeric(thetas[r, ])
    out$lp_j0 <- c(out$lp_j0, fit$log_prob(unconstrained_variables = th, jacobian = FALSE))
    out$lp_j1 <- c(out$lp_j1, fit$log_prob(unconstrained_variables = th, jacobian = TRUE))
    out$grad_j0[[r]] <- as.numeric(fit$grad_log_prob(unconstrained_variables = th, jacobian = FALSE))
    out$grad_j1[[r]] <- as.numeric(fit$grad_log_prob(unconstrained_variables = th, jacobian = TRUE))
  }
  write_json(out, file.path(kit, paste0(name, ".cmdstanr.json")), digits = NA, auto_unbox = TRUE)
  message(sprintf("%s (%s): %d vectors written", name, prog, nrow(thetas)))
}

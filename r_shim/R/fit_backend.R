# fit_backend.R -- what replaces lines 290-432 of R/findSpatiallyVariableFeaturesBayes.R in the
# reference when the B200 engine is used.  Everything before (:99-288: argument checks, coordinate
# scaling, expression extraction, k-means centroids, length-scale) and after (:436-508: write-back
# into the Seurat / SpatialExperiment object) stays exactly as it is.  NOT exercised in this
# repository's image (no R); the Python mirror bayesvg_b200/api.py runs the same sequence of C-ABI
# calls and is what the tests drive.
#
# in scope at :290  : spatial_mtx (M x 2, scaled), kmeans_centers (k x 2), lscale, expr_df (long,
#                     gene-major), gene_mapping, likelihood, kernel, kernel.smoothness, kernel.period,
#                     n.iter, mle.init, mle.iter, n.draws, elbo.samples, random.seed, verbose;
#                     gene_depths (:303, log(rowSums(counts[naive.hvgs, ]))) when gene.depth.adjust = TRUE, else NULL
bvg_fit_backend <- function(spatial_mtx, kmeans_centers, lscale, expr_df, gene_mapping, likelihood, kernel,
                            kernel.smoothness, kernel.period, n.iter, mle.init, mle.iter, n.draws,
                            elbo.samples, random.seed, verbose, gene_depths = NULL, device = 0L, save.model = FALSE,
                            algorithm = "meanfield", n.cores = 2L) {
  M <- nrow(spatial_mtx)
  G <- length(unique(expr_df$gene))
  kernel_id <- match(kernel, c("exp_quad", "matern", "periodic")) - 1L
  # :265-288  phi <- kernel(...); phi_ortho <- qr.Q(qr(phi, LAPACK = TRUE))
  phi_ortho <- .Call("R_bvg_basis_build", spatial_mtx, kmeans_centers, as.double(lscale), kernel_id,
                     as.double(kernel.smoothness), as.double(kernel.period), as.integer(device))
  # :320-327  data_list: y is the gene-major long vector, i.e. an M x G column-major matrix
  y <- expr_df$gene_expression
  dim(y) <- c(M, G)
  # :337-364  cmdstan_model / compile / optimize / variational  ->  one handle, one run
  fit <- .Call("R_bvg_fit_create", list(likelihood = ifelse(likelihood == "gaussian", 0L, 1L), n_iter = as.integer(n.iter),
                                        mle_init = as.integer(mle.init), mle_iter = as.integer(mle.iter),
                                        n_draws = as.integer(n.draws), elbo_samples = as.integer(elbo.samples),
                                        seed = as.integer(random.seed), verbose = as.integer(verbose),
                                        depth_adjust = as.integer(!is.null(gene_depths)),  # approxGP2 / approxGP3 (:293-318)
                                        algorithm = match(algorithm, c("meanfield", "fullrank", "pathfinder")) - 1L,  # :357-375
                                        pf_num_paths = as.integer(n.cores),                # num_paths = n.cores (:372)
                                        device = as.integer(device)))
  on.exit(.Call("R_bvg_fit_destroy", fit), add = TRUE)
  .Call("R_bvg_fit_set_phi", fit, phi_ortho)
  if (!is.null(gene_depths)) .Call("R_bvg_fit_set_gene_depths", fit, as.double(gene_depths))
  .Call("R_bvg_fit_set_y", fit, y)
  .Call("R_bvg_fit_run", fit)
  # :419-432  fit_vi$summary(variables = "amplitude") %>% rename / join / dispersion / rank
  s <- .Call("R_bvg_fit_summary", fit, G)
  amplitude_summary <- data.frame(gene_id = as.character(seq_len(G)),
                                  amplitude_mean = s[, 1], amplitude_median = s[, 2], amplitude_sd = s[, 3],
                                  amplitude_mad = s[, 4], amplitude_ci_ll = s[, 5], amplitude_ci_ul = s[, 6]) %>%
                       dplyr::inner_join(gene_mapping, by = "gene_id") %>%
                       dplyr::relocate(gene) %>%
                       dplyr::select(-gene_id) %>%
                       dplyr::mutate(amplitude_dispersion = amplitude_sd^2 / amplitude_mean) %>%
                       dplyr::arrange(dplyr::desc(amplitude_mean)) %>%
                       dplyr::mutate(amplitude_mean_rank = dplyr::row_number()) %>%
                       as.data.frame() %>%
                       magrittr::set_rownames(.$gene)
  # :483-489  save.model = TRUE: the draws of the whole model as a posterior::draws_matrix (what fit_vi$draws() holds);
  #           column names are built as in bayesvg_b200/engine.py::Fit.posterior_columns
  if (save.model) attr(amplitude_summary, "draws") <- .Call("R_bvg_fit_posterior", fit, as.integer(n.draws))
  amplitude_summary
}


# What replaces lines 139-291 of R/clusterSVGsBayes.R (compile, $variational, the foreach responsibility loop, the log_lik summary).
# in scope at :139 : svg_mtx_pca (irlba::prcomp_irlba result), svgs, n.clust, algorithm, n.iter, n.draws, elbo.samples, random.seed
bvg_cluster_backend <- function(svg_mtx_pca, svgs, n.clust, algorithm, n.iter, n.draws, elbo.samples, random.seed, device = 0L) {
  X <- svg_mtx_pca$x
  N <- nrow(X); D <- ncol(X); K <- as.integer(n.clust)
  opts <- as.double(c(n.iter, n.draws, elbo.samples, 100, 50, 0, 0.01))       # eval_elbo, adapt_iter, eta (adaptive), tol_rel_obj: CmdStan defaults
  res <- .Call("R_bvg_gmm_fit", X, K, algorithm == "fullrank", as.integer(random.seed), opts, as.integer(device))
  soft_assignments <- t(res[[1]])                                              # :241-254
  cluster_df <- as.data.frame(soft_assignments)
  colnames(cluster_df) <- paste0("prob_cluster_", seq(K))
  cluster_df <- dplyr::mutate(cluster_df, gene = svgs, .before = 1)
  cluster_df$assigned_cluster <- max.col(soft_assignments, ties.method = "first")  # :262-266 which.max
  log_likelihood <- res[[3]][5]                                                # :281-284
  p <- (K - 1) + 2 * K * D                                                     # :286
  list(pca_embedding = X, cluster_df = cluster_df, model_fit = t(res[[2]]), log_likelihood = log_likelihood,
       BIC = -2 * log_likelihood + p * log(N))
}

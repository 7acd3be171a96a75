/* bvg_rshim.c -- logic-free .Call shim between R and libbayesvg_b200.so (include/bayesvg_b200.h).
 * NOT compiled in this repository's image (no R, no Rinternals.h); see INTEGRATION.md.
 * Build inside an R package:  PKG_CPPFLAGS = -I<repo>/include ; PKG_LIBS = -L<libdir> -lbayesvg_b200
 * Every entry point converts SEXP <-> plain pointers, calls one bvg_* function and turns a non-zero
 * status into an R error carrying bvg_last_error().  The fit handle is an external pointer with a
 * finalizer, so an Rf_error long-jump cannot leak device memory. */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>
#include "bayesvg_b200.h"

static void chk(int rc) { if (rc != BVG_OK) Rf_error("bayesvg_b200: %s", bvg_last_error()); }
static void fit_finalizer(SEXP p) {
  bvg_fit *f = (bvg_fit *)R_ExternalPtrAddr(p);
  if (f) { bvg_fit_destroy(f); R_ClearExternalPtr(p); }
}
static bvg_fit *get_fit(SEXP p) {
  bvg_fit *f = (bvg_fit *)R_ExternalPtrAddr(p);
  if (!f) Rf_error("bayesvg_b200: fit handle already released");
  return f;
}
/* numeric input as a double vector / matrix: integer (or logical) storage is coerced, anything else is an error.
 * The result must be PROTECTed by the caller (it may be a fresh allocation). */
static SEXP as_real(SEXP x, const char *what) {
  if (TYPEOF(x) == REALSXP) return x;
  if (TYPEOF(x) == INTSXP || TYPEOF(x) == LGLSXP) return Rf_coerceVector(x, REALSXP);
  Rf_error("bayesvg_b200: %s must be numeric", what);
  return R_NilValue;
}
/* sizes come from the handle, never from the caller: a wrong G or D on the R side must not become a heap overflow */
static void fit_shape(bvg_fit *f, int64_t *M, int64_t *G, int *k, int64_t *D) { chk(bvg_fit_shape(f, M, G, k, D)); }

/* .Call("R_bvg_basis_build", coords (M x 2), centers (k x 2), lscale, kernel_id, nu, period, device) */
SEXP R_bvg_basis_build(SEXP coords, SEXP centers, SEXP lscale, SEXP kernel, SEXP nu, SEXP period, SEXP device) {
  coords = PROTECT(as_real(coords, "coords"));
  centers = PROTECT(as_real(centers, "centers"));
  if (Rf_ncols(coords) != 2 || Rf_ncols(centers) != 2) Rf_error("bayesvg_b200: coords and centers must have two columns");
  const R_xlen_t M = Rf_nrows(coords);
  const int k = Rf_nrows(centers);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)M, k));
  chk(bvg_basis_build(Rf_asInteger(device), REAL(coords), (int64_t)M, REAL(centers), k, Rf_asReal(lscale),
                      Rf_asInteger(kernel), Rf_asReal(nu), Rf_asReal(period), REAL(out)));
  UNPROTECT(3);
  return out;
}

/* .Call("R_bvg_variogram_lscale", spatial_mtx (M x 2, scaled), expr_mtx[naive.hvgs, ] (G x M), kernel.smoothness, device)
 * -> list(lscale =, ranges =): replaces the foreach / gstat loop of :208-260 */
SEXP R_bvg_variogram_lscale(SEXP coords, SEXP expr, SEXP kappa, SEXP device) {
  coords = PROTECT(as_real(coords, "coords"));
  expr = PROTECT(as_real(expr, "expr"));
  const int64_t M = Rf_nrows(coords), G = Rf_nrows(expr);
  if (Rf_ncols(coords) != 2 || Rf_ncols(expr) != M) Rf_error("bayesvg_b200: coords must be M x 2 and expr G x M");
  SEXP ranges = PROTECT(Rf_allocVector(REALSXP, G));
  double ls = 0;
  chk(bvg_variogram_lscale(Rf_asInteger(device), REAL(coords), M, REAL(expr), G, Rf_asReal(kappa), REAL(ranges), &ls));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, Rf_ScalarReal(ls));
  SET_VECTOR_ELT(out, 1, ranges);
  UNPROTECT(4);
  return out;
}
/* .Call("R_bvg_gmm_fit", svg_mtx_pca$x (N x D), K, fullrank, seed, opts (7 doubles), device)
 * -> list(soft = N x K, draws = n_draws x (K + 2 K D), stats = 8 doubles): replaces R/clusterSVGsBayes.R:139-291 */
SEXP R_bvg_gmm_fit(SEXP X, SEXP K, SEXP fullrank, SEXP seed, SEXP opts, SEXP device) {
  X = PROTECT(as_real(X, "X"));
  opts = PROTECT(as_real(opts, "opts"));
  if (Rf_xlength(opts) != 7) Rf_error("bayesvg_b200: opts must hold 7 numbers");
  const int64_t N = Rf_nrows(X);
  const int D = Rf_ncols(X), k = Rf_asInteger(K), nd = (int)REAL(opts)[1];
  if (k < 1 || nd < 1) Rf_error("bayesvg_b200: K and n_draws must be positive");
  SEXP softT = PROTECT(Rf_allocMatrix(REALSXP, k, (int)N));                 /* the ABI writes [N][K] row-major = K x N column-major */
  SEXP drawsT = PROTECT(Rf_allocMatrix(REALSXP, k + 2 * k * D, nd));
  SEXP stats = PROTECT(Rf_allocVector(REALSXP, 8));
  chk(bvg_gmm_fit(Rf_asInteger(device), REAL(X), N, D, k, Rf_asLogical(fullrank), (uint64_t)Rf_asInteger(seed), REAL(opts), NULL, NULL,
                  REAL(drawsT), REAL(softT), REAL(stats)));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, softT);    /* t() on the R side */
  SET_VECTOR_ELT(out, 1, drawsT);
  SET_VECTOR_ELT(out, 2, stats);
  UNPROTECT(6);
  return out;
}
/* .Call("R_bvg_kmeans", spatial_mtx, k, iter.max, nstart, seed, device) -> k x 2 centres (:200-204) */
SEXP R_bvg_kmeans(SEXP coords, SEXP k, SEXP iter_max, SEXP nstart, SEXP seed, SEXP device) {
  coords = PROTECT(as_real(coords, "coords"));
  const int kk = Rf_asInteger(k);
  if (Rf_ncols(coords) != 2 || kk < 1) Rf_error("bayesvg_b200: coords must be M x 2 and k positive");
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, kk, 2));
  chk(bvg_kmeans(Rf_asInteger(device), REAL(coords), (int64_t)Rf_nrows(coords), kk, Rf_asInteger(iter_max), Rf_asInteger(nstart),
                 (uint64_t)Rf_asInteger(seed), REAL(out), NULL, NULL));
  UNPROTECT(2);
  return out;
}

/* .Call("R_bvg_fit_create", list(likelihood=, precision=, n_iter=, mle_init=, mle_iter=, n_draws=,
 *        elbo_samples=, seed=, verbose=, device=, depth_adjust=)) -> external pointer */
static int geti(SEXP lst, const char *name, int dflt) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  for (R_xlen_t i = 0; i < Rf_xlength(lst); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return Rf_asInteger(VECTOR_ELT(lst, i));
  return dflt;
}
SEXP R_bvg_fit_create(SEXP opts) {
  bvg_config cfg;
  bvg_config_default(&cfg);
  cfg.likelihood = geti(opts, "likelihood", cfg.likelihood);
  cfg.precision = geti(opts, "precision", cfg.precision);
  cfg.engine = geti(opts, "engine", cfg.engine);
  cfg.device = geti(opts, "device", cfg.device);
  cfg.n_iter = geti(opts, "n_iter", cfg.n_iter);
  cfg.mle_init = geti(opts, "mle_init", cfg.mle_init);
  cfg.mle_iter = geti(opts, "mle_iter", cfg.mle_iter);
  cfg.n_draws = geti(opts, "n_draws", cfg.n_draws);
  cfg.elbo_samples = geti(opts, "elbo_samples", cfg.elbo_samples);
  cfg.verbose = geti(opts, "verbose", cfg.verbose);
  cfg.depth_adjust = geti(opts, "depth_adjust", cfg.depth_adjust);
  cfg.algorithm = geti(opts, "algorithm", cfg.algorithm);            /* 0 meanfield, 1 fullrank, 2 pathfinder (:108) */
  cfg.pf_num_paths = geti(opts, "pf_num_paths", cfg.pf_num_paths);   /* num_paths = n.cores (:372) */
  cfg.lbfgs_history = geti(opts, "lbfgs_history", cfg.lbfgs_history);
  cfg.seed = (uint64_t)geti(opts, "seed", (int)cfg.seed);
  bvg_fit *f = NULL;
  chk(bvg_fit_create(&cfg, &f));
  SEXP p = PROTECT(R_MakeExternalPtr(f, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(p, fit_finalizer, TRUE);
  UNPROTECT(1);
  return p;
}
SEXP R_bvg_fit_set_phi(SEXP p, SEXP phi) {
  phi = PROTECT(as_real(phi, "phi"));
  chk(bvg_fit_set_phi(get_fit(p), REAL(phi), (int64_t)Rf_nrows(phi), Rf_ncols(phi)));
  UNPROTECT(1);
  return R_NilValue;
}
/* y: M x G matrix (= the gene-major long vector of :176-183 with dim(y) <- c(M, G)); double or integer */
SEXP R_bvg_fit_set_y(SEXP p, SEXP y) {
  const int64_t M = Rf_nrows(y), G = Rf_ncols(y);
  if (TYPEOF(y) == INTSXP) chk(bvg_fit_set_y_i32(get_fit(p), INTEGER(y), M, G));
  else if (TYPEOF(y) == REALSXP) chk(bvg_fit_set_y_f64(get_fit(p), REAL(y), M, G));
  else Rf_error("bayesvg_b200: y must be a double or integer matrix");
  return R_NilValue;
}
/* gene_depths: log(rowSums(counts[naive.hvgs, ])), findSpatiallyVariableFeaturesBayes.R:303 (gene.depth.adjust = TRUE) */
SEXP R_bvg_fit_set_gene_depths(SEXP p, SEXP d) {
  d = PROTECT(as_real(d, "gene_depths"));
  chk(bvg_fit_set_gene_depths(get_fit(p), REAL(d), (int64_t)Rf_xlength(d)));
  UNPROTECT(1);
  return R_NilValue;
}
SEXP R_bvg_fit_run(SEXP p) { chk(bvg_fit_run(get_fit(p))); return R_NilValue; }
/* -> G x 8 matrix: mean, median, sd, mad, q5, q95, dispersion, rank.  G is read from the handle; the second argument is
 * kept for call compatibility and only checked. */
SEXP R_bvg_fit_summary(SEXP p, SEXP G) {
  int64_t Gh = 0;
  fit_shape(get_fit(p), NULL, &Gh, NULL, NULL);
  if (!Rf_isNull(G) && (int64_t)Rf_asInteger(G) != Gh) Rf_error("bayesvg_b200: the fit holds %lld genes, not %d", (long long)Gh, Rf_asInteger(G));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)Gh, 8));
  chk(bvg_fit_get_summary(get_fit(p), REAL(out)));
  UNPROTECT(1);
  return out;
}
/* the parity hook: list(lp =, grad =) at an unconstrained vector, like fit$log_prob / fit$grad_log_prob */
SEXP R_bvg_log_prob(SEXP p, SEXP theta, SEXP jacobian, SEXP propto) {
  int64_t D = 0;
  fit_shape(get_fit(p), NULL, NULL, NULL, &D);
  theta = PROTECT(as_real(theta, "theta"));
  if ((int64_t)Rf_xlength(theta) != D) Rf_error("bayesvg_b200: theta has %lld entries, the model has %lld unconstrained parameters", (long long)Rf_xlength(theta), (long long)D);
  SEXP grad = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)D));
  double lp = 0;
  chk(bvg_log_prob(get_fit(p), REAL(theta), Rf_asLogical(jacobian), Rf_asLogical(propto), &lp, REAL(grad)));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, Rf_ScalarReal(lp));
  SET_VECTOR_ELT(out, 1, grad);
  UNPROTECT(3);
  return out;
}
/* D is read from the handle; the second argument is kept for call compatibility and only checked */
SEXP R_bvg_fit_variational(SEXP p, SEXP D) {
  int64_t Dh = 0;
  fit_shape(get_fit(p), NULL, NULL, NULL, &Dh);
  if (!Rf_isNull(D) && (int64_t)Rf_asInteger(D) != Dh) Rf_error("bayesvg_b200: the model has %lld unconstrained parameters, not %d", (long long)Dh, Rf_asInteger(D));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)Dh, 2));
  chk(bvg_fit_get_variational(get_fit(p), REAL(out), REAL(out) + Dh));
  UNPROTECT(1);
  return out;
}
/* save.model = TRUE (:483-489): the n x ncol draws matrix behind cmdstanr's fit_vi$draws(); columns lp__, log_p__,
 * log_g__, the parameters constrained in declaration order, beta0[G] (hierarchical intercept), amplitude_sq[G] */
SEXP R_bvg_fit_posterior(SEXP p, SEXP n) {
  int64_t ncol = 0;
  chk(bvg_fit_posterior_ncol(get_fit(p), &ncol));
  if (Rf_asInteger(n) < 1) Rf_error("bayesvg_b200: n must be positive");
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, Rf_asInteger(n), (int)ncol));
  chk(bvg_fit_get_posterior(get_fit(p), Rf_asInteger(n), REAL(out)));
  UNPROTECT(1);
  return out;
}
SEXP R_bvg_fit_destroy(SEXP p) { fit_finalizer(p); return R_NilValue; }
/* .onUnload: hand the device blocks bvg_fit_destroy keeps for the next fit back to the driver */
SEXP R_bvg_release_cached_memory(void) { bvg_release_cached_memory(); return R_NilValue; }

static const R_CallMethodDef calls[] = {
  {"R_bvg_basis_build", (DL_FUNC)&R_bvg_basis_build, 7}, {"R_bvg_fit_create", (DL_FUNC)&R_bvg_fit_create, 1},
  {"R_bvg_fit_set_phi", (DL_FUNC)&R_bvg_fit_set_phi, 2}, {"R_bvg_fit_set_y", (DL_FUNC)&R_bvg_fit_set_y, 2},
  {"R_bvg_fit_run", (DL_FUNC)&R_bvg_fit_run, 1},         {"R_bvg_fit_summary", (DL_FUNC)&R_bvg_fit_summary, 2},
  {"R_bvg_log_prob", (DL_FUNC)&R_bvg_log_prob, 4},       {"R_bvg_fit_variational", (DL_FUNC)&R_bvg_fit_variational, 2},
  {"R_bvg_fit_destroy", (DL_FUNC)&R_bvg_fit_destroy, 1}, {"R_bvg_fit_set_gene_depths", (DL_FUNC)&R_bvg_fit_set_gene_depths, 2},
  {"R_bvg_fit_posterior", (DL_FUNC)&R_bvg_fit_posterior, 2},
  {"R_bvg_variogram_lscale", (DL_FUNC)&R_bvg_variogram_lscale, 4}, {"R_bvg_kmeans", (DL_FUNC)&R_bvg_kmeans, 6},
  {"R_bvg_gmm_fit", (DL_FUNC)&R_bvg_gmm_fit, 6},         {"R_bvg_release_cached_memory", (DL_FUNC)&R_bvg_release_cached_memory, 0},
  {NULL, NULL, 0}};
void R_init_bayesVGb200(DllInfo *dll) {
  R_registerRoutines(dll, NULL, calls, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}

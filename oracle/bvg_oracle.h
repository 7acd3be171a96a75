/*
 * bvg_oracle.h -- CPU oracle for the bayesVG SVG hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the reference's arithmetic for this path lives in CmdStan / Stan Math
 * (un-vendored, version-unpinned: DESCRIPTION:58 "SystemRequirements: CmdStan") and the
 * reference's own tests hold no numeric golden vectors (tests/testthat/test_bayesVG.R:144-219
 * assert classes and shapes only).  Neither R nor CmdStan exists in this image, so this file is a
 * restatement of inst/approxGP{,2,3,4}.stan and inst/clusterSVGs.stan and of the published Stan algorithms
 * (meanfield and fullrank ADVI, L-BFGS, Pathfinder with PSIS resampling), pinned only against independent
 * literal transcriptions of the Stan programs (tests/golden/make_golden.py, tests/test_oracle_golden.py:
 * scipy.stats log-densities + torch-autograd fp64 / central differences), the BFGS recursion (Pathfinder's
 * approximations) and brute-force loops (k-means fixed points, empirical variograms in oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (bayesvg_b200/) never does.
 */
#ifndef BVG_ORACLE_H
#define BVG_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  BVGO_GAUSSIAN = 0 /* inst/approxGP.stan */, BVGO_NB = 1 /* inst/approxGP4.stan */,
  BVGO_GAUSSIAN_DEPTH = 2 /* inst/approxGP2.stan */, BVGO_NB_DEPTH = 3 /* inst/approxGP3.stan */,
  BVGO_GMM = 4 /* inst/clusterSVGs.stan (bvgo_gmm_create) */
};
enum { BVGO_KERNEL_EXP_QUAD = 0, BVGO_KERNEL_MATERN = 1, BVGO_KERNEL_PERIODIC = 2 };

typedef struct bvgo_ctx bvgo_ctx;

/* y: M x G column-major (R's gene-major long vector, findSpatiallyVariableFeaturesBayes.R:176-183),
 * phi: M x k column-major (inst/approxGP.stan:8).  Buffers are borrowed, not copied. */
bvgo_ctx *bvgo_create(int model, int64_t M, int64_t G, int k, const double *phi, const double *y,
                      int nthreads);
void bvgo_destroy(bvgo_ctx *);
int64_t bvgo_dim(int model, int64_t G, int k);
/* depth-adjusted models: gene_depths[G] (log of the gene's total count, R/findSpatiallyVariableFeaturesBayes.R:303);
 * the pointer must outlive the context */
void bvgo_set_depths(bvgo_ctx *c, const double *gene_depths);

/* log density at an unconstrained vector (Stan declaration order);
 * grad may be NULL (forward only).  Mirrors log_prob<propto, jacobian>. */
double bvgo_log_prob(bvgo_ctx *, const double *theta, int jacobian, int propto, double *grad);

/* counter-based RNG shared (by specification, not by code) with the CUDA engine */
void bvgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double bvgo_normal(uint64_t seed, uint32_t stream, uint32_t t, uint64_t i);

typedef struct {
  int iter;          /* n.iter   = 3000  (findSpatiallyVariableFeaturesBayes.R:82)  */
  int elbo_samples;  /* 150 (:110-116) */
  int eval_elbo;     /* CmdStan default 100 */
  int adapt_iter;    /* CmdStan default 50  */
  double tol_rel_obj;/* CmdStan default 0.01 */
  double eta;        /* <=0: adapt (CmdStan default) */
  uint64_t seed;     /* random.seed = 312 (:95) */
} bvgo_advi_cfg;

typedef struct {
  double eta;        /* chosen step size */
  int iters;         /* main-loop iterations executed */
  int grad_evals;    /* calc_ELBO_grad calls (adapt + main) */
  int elbo_evals;    /* calc_ELBO calls */
  int converged;     /* 1 mean, 2 median, 3 both, 0 max-iter */
  double elbo;       /* last ELBO */
  int n_trace;       /* rows written to trace */
} bvgo_advi_result;

/* meanfield ADVI started at theta_init (mu = theta_init, omega = 0).
 * trace: optional [iter/eval_elbo][4] = (iter, elbo, delta_mean, delta_med). */
int bvgo_advi(bvgo_ctx *, const double *theta_init, const bvgo_advi_cfg *, double *mu, double *omega,
              bvgo_advi_result *, double *trace, int trace_cap);

/* single ADVI pieces exposed so the CUDA engine can be checked step by step */
double bvgo_elbo(bvgo_ctx *, const double *mu, const double *omega, int S, uint64_t seed,
                 uint32_t t0);
/* one calc_ELBO_grad + step; hist is the running s (2D); t = gradient draw counter (global),
 * it = 1-based iteration counter inside the current phase */
int bvgo_advi_step(bvgo_ctx *, double *mu, double *omega, double *hist, double eta, int it,
                   uint64_t seed, uint32_t t);

/* algorithm = "fullrank" [U: stan::variational::normal_fullrank]: q = N(mu, L L^T), L lower triangular, PACKED row-major
 * (row i starts at i (i + 1) / 2; D (D + 1) / 2 entries), initialised to the identity.  hist = [D | D (D + 1) / 2]. */
int bvgo_advi_fullrank(bvgo_ctx *, const double *theta_init, const bvgo_advi_cfg *, double *mu, double *L_packed,
                       bvgo_advi_result *, double *trace, int trace_cap);
double bvgo_fr_elbo(bvgo_ctx *, const double *mu, const double *L_packed, int S, uint64_t seed, uint32_t t0);
int bvgo_fr_step(bvgo_ctx *, double *mu, double *L_packed, double *hist, double eta, int it, uint64_t seed, uint32_t t);
int bvgo_summary_fullrank(int model, int64_t G, int k, const double *mu, const double *L_packed, int n_draws,
                          uint64_t seed, double *out);

typedef struct {
  int max_iter;      /* mle.iter = 1000 (:89) */
  int history;       /* 25 (:353) */
  double init_alpha; /* 1e-3 */
  double tol_obj, tol_rel_obj, tol_grad, tol_rel_grad, tol_param; /* CmdStan defaults */
} bvgo_lbfgs_cfg;
typedef struct { int iters; int fevals; int term; double lp; } bvgo_lbfgs_result;
int bvgo_lbfgs(bvgo_ctx *, const double *theta0, const bvgo_lbfgs_cfg *, double *theta_out,
               bvgo_lbfgs_result *);

/* algorithm = "pathfinder" (:366-375) [U: stan::services::pathfinder]; see bvg_oracle.c */
typedef struct {
  int max_lbfgs_iters;  /* 200 (:373) */
  int history;          /* history_size = 25 (:375); <= 25 */
  int num_elbo_draws;   /* elbo.samples = 50 for pathfinder (:111-112) */
  int num_draws;        /* single-path draws, CmdStan default 1000 */
  int num_paths;        /* n.cores (:372) */
  int num_psis_draws;   /* draws = n.draws (:370) */
  double init_radius;   /* init = 0.01 (:368) */
  uint64_t seed;        /* random.seed */
} bvgo_pf_cfg;
typedef struct { int lbfgs_iters, lbfgs_term, best_iter, n_pairs; double best_elbo; } bvgo_pf_info;
int bvgo_pf_single(bvgo_ctx *, const bvgo_pf_cfg *, int path, double *amp /*[num_draws][G]*/, double *lp_ratio /*[num_draws]*/,
                   bvgo_pf_info *, double *elbos /* optional [max_lbfgs_iters + 1] */);
int bvgo_pf_single_phi(bvgo_ctx *, const bvgo_pf_cfg *, int path, double *phi /*[num_draws][D] unconstrained*/, double *lp_ratio,
                       bvgo_pf_info *, double *elbos);
double bvgo_psis_weights(const double *log_ratios, int S, double *weights); /* returns the Pareto k estimate */
int bvgo_pathfinder(bvgo_ctx *, const bvgo_pf_cfg *, double *summary /*G x 8*/, bvgo_pf_info *infos /*[num_paths]*/,
                    int *idx /* optional [num_psis_draws] */, double *khat);
int bvgo_pf_approx_test(int64_t D, int n, const double *alpha, const double *S, const double *Y, const double *x,
                        const double *g, int nu, const double *u, double *phi, double *logq, double *xc);
double bvgo_uniform(uint64_t seed, uint32_t stream, uint32_t t, uint64_t i);

/* inst/clusterSVGs.stan (R/clusterSVGsBayes.R:139-291): the mixture model over the PCA embedding X [N][D] row-major.  The
 * returned context works with bvgo_log_prob / bvgo_advi / bvgo_advi_fullrank (dimension K - 1 + 2 K D). */
bvgo_ctx *bvgo_gmm_create(const double *X, int64_t N, int D, int K);
int bvgo_gmm_post(bvgo_ctx *, const double *mu, const double *var /* omega or packed L */, int fullrank, int n_draws,
                  uint64_t seed, double *draws /* [n_draws][K + 2 K D] or NULL */, double *soft /* [N][K] */, double *loglik);

/* amplitude summaries from n_draws draws of exp(mu_u + exp(omega_u) eps):
 * out G x 8 col-major: mean, median, sd, mad, q5, q95, dispersion, rank (1 = largest mean) */
int bvgo_summary(int model, int64_t G, int k, const double *mu, const double *omega, int n_draws,
                 uint64_t seed, int64_t gene_offset, int64_t G_total, double *out);

/* basis builder (findSpatiallyVariableFeaturesBayes.R:265-288): kernel at centroids + pivoted
 * Householder QR (dgeqp3-equivalent), thin Q out (M x k col-major).  raw_out optional. */
int bvgo_basis(const double *coords /*M x 2 col-major*/, int64_t M, const double *centers /*k x 2*/,
               int k, double lscale, int kernel, double nu, double period, double *phi_out,
               double *raw_out);

/* centroids (findSpatiallyVariableFeaturesBayes.R:200-204, stats::kmeans(centers = k, iter.max, nstart)): Lloyd-Forgy
 * from nstart Philox-drawn sets of k distinct rows, fixed-point sums (serial restatement of csrc/k_kmeans.cu; R's own
 * Hartigan-Wong + ambient RNG is not reproducible).  info = {best start, iterations, converged}. */
int bvgo_kmeans(const double *coords /*M x 2 col-major*/, int64_t M, int k, int iter_max, int nstart,
                uint64_t seed, double *centers_out /*k x 2*/, double *tot_withinss, int32_t *info);

double bvgo_digamma(double x);

#ifdef __cplusplus
}
#endif
#endif

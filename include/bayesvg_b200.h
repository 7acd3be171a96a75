/*
 * bayesvg_b200.h -- C ABI of the B200-native engine for bayesVG's SVG hot path.
 *
 * The reference has no FFI: its only seam for this path is the sequence of cmdstanr method calls
 * in R/findSpatiallyVariableFeaturesBayes.R:337-364 and :419 (paths relative to the reference):
 *     cmdstan_model() / $compile()      :337-341   -> nothing (kernels are precompiled)
 *     data_list                          :320-327   -> bvg_fit_set_phi / bvg_fit_set_y_*
 *     mod$optimize(init=0, jacobian=F)   :346-353   -> bvg_fit_mle            (or bvg_fit_run)
 *     mod$variational(meanfield, ...)    :357-364   -> bvg_fit_advi           (or bvg_fit_run)
 *     fit_vi$summary("amplitude")        :419-432   -> bvg_fit_summarize + bvg_fit_get_summary
 *     fit$log_prob / fit$grad_log_prob   (cmdstanr) -> bvg_log_prob           (the parity hook)
 *     phi <- kernel(...); qr.Q(qr(phi))  :265-288   -> bvg_basis_build
 * and, for the rows either side of and next to that path:
 *     stats::kmeans(centers = k, ...)    :200-204   -> bvg_kmeans
 *     lscale.estimator = "variogram"     :208-260   -> bvg_variogram_lscale
 *     mod$variational("fullrank")        :357-364   -> cfg.algorithm = BVG_ALG_FULLRANK   (same bvg_fit_run)
 *     mod$pathfinder(...)                :366-375   -> cfg.algorithm = BVG_ALG_PATHFINDER (same bvg_fit_run)
 *     save.model = TRUE / extractModel() :483-489   -> bvg_fit_get_posterior
 *     clusterSVGsBayes()   R/clusterSVGsBayes.R:139-291 -> bvg_gmm_fit, bvg_gmm_log_prob
 * An R package binds these through a logic-free .Call shim.
 *
 * Conventions: every matrix is COLUMN-MAJOR like R.  All pointers are HOST pointers owned by the
 * caller; the library copies on set_* and writes results into caller-allocated buffers.  Every
 * function returns 0 on success or a negative bvg_status; the message is in bvg_last_error()
 * (thread-local).  Nothing throws across the ABI.  There is no CPU fallback: without a CUDA
 * device every compute entry point fails with BVG_ERR_CUDA.
 */
#ifndef BAYESVG_B200_H
#define BAYESVG_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BVG_ABI_VERSION 2

typedef enum {
  BVG_OK = 0,
  BVG_ERR_ARG = -1,      /* invalid argument / unsupported configuration */
  BVG_ERR_CUDA = -2,     /* CUDA runtime or driver error, or no device */
  BVG_ERR_STATE = -3,    /* call order (e.g. run before set_y) */
  BVG_ERR_NUMERIC = -4,  /* non-finite log density where Stan would raise */
  BVG_ERR_COMM = -5      /* NCCL error */
} bvg_status;

/* likelihood x gene.depth.adjust selects the Stan program (R/findSpatiallyVariableFeaturesBayes.R:293-333):
 *   gaussian: inst/approxGP.stan, depth-adjusted inst/approxGP2.stan;  nb: inst/approxGP4.stan, depth-adjusted inst/approxGP3.stan */
typedef enum { BVG_GAUSSIAN = 0, BVG_NB = 1 } bvg_likelihood;
typedef enum { BVG_KERNEL_EXP_QUAD = 0, BVG_KERNEL_MATERN = 1, BVG_KERNEL_PERIODIC = 2 } bvg_kernel;
typedef enum {
  BVG_PREC_FP64 = 0, /* Y and all arithmetic fp64: the <=1e-6 parity path */
  BVG_PREC_FAST = 1  /* Y fp32, likelihood contraction fp32 / 3xTF32; parameters stay fp64 */
} bvg_precision;
typedef enum {
  BVG_ENGINE_AUTO = 0,    /* tcgen05 when precision = FAST (fused kernel for k+1 <= 24, two-GEMM kernels above), else SIMT */
  BVG_ENGINE_SIMT = 1,    /* CUDA-core kernel (fp64 or fp32) */
  BVG_ENGINE_TCGEN05 = 2  /* TMA + tcgen05/TMEM kernel (FAST only) */
} bvg_engine;

/* Mirrors the arguments of findSpatiallyVariableFeaturesBayes() (:78-97) that reach the fit, plus
 * the CmdStan defaults the reference leaves untouched. */
typedef struct {
  int32_t abi_version;   /* BVG_ABI_VERSION */
  int32_t likelihood;    /* bvg_likelihood   (likelihood = "gaussian" | "nb", :80) */
  int32_t precision;     /* bvg_precision */
  int32_t engine;        /* bvg_engine */
  int32_t device;        /* CUDA ordinal */
  int32_t n_iter;        /* n.iter = 3000 (:82) */
  int32_t mle_init;      /* mle.init = TRUE (:88) */
  int32_t mle_iter;      /* mle.iter = 1000 (:89) */
  int32_t n_draws;       /* n.draws = 1000 (:91) */
  int32_t elbo_samples;  /* elbo.samples = 150 for meanfield (:110-116) */
  int32_t eval_elbo;     /* CmdStan default 100 */
  int32_t adapt_iter;    /* CmdStan default 50 */
  int32_t lbfgs_history; /* history_size = 25 (:353) */
  int32_t verbose;       /* print the ELBO table to stderr */
  double tol_rel_obj;    /* CmdStan default 0.01 */
  double eta;            /* <= 0: adapt over {100,10,1,0.1,0.01} (CmdStan default) */
  uint64_t seed;         /* random.seed = 312 (:95) */
  /* gene sharding: this process holds genes [gene_offset, gene_offset + G) of G_total */
  int32_t rank, world_size;
  int64_t gene_offset, G_total; /* G_total = 0 -> unsharded */
  /* default 1: where they exist (gaussian likelihood, hierarchical intercept, n.basis.fns <= 30) the forward-only ELBO
   * estimates (150 draws every 100 iterations) are evaluated from per-gene sufficient statistics (sum y^2, Phi'^T y,
   * Phi'^T Phi', fp64, built once on the device) instead of 150 dense passes over Y: the same draws and the same
   * value up to round-off in O(k^2 G) per draw; every other model, and 0, takes the dense batched
   * kernel (k_elbo_persist).  The gradient path is always the dense kernel. */
  int32_t elbo_suffstat;
  /* gene.depth.adjust (:85): intercept beta0 + beta1 * gene_depths[g] (approxGP2/3.stan) instead of the
   * hierarchical per-gene intercept; needs bvg_fit_set_gene_depths */
  int32_t depth_adjust;
  /* algorithm (:87, :108): BVG_ALG_MEANFIELD (default); BVG_ALG_FULLRANK -- q = N(mu, L L^T) with the packed
   * lower-triangular L and its step-size history resident in HBM (D (D + 1) / 2 * 16 bytes: 35 GB at 3,000 genes,
   * k = 20); BVG_ALG_PATHFINDER -- multi-path Pathfinder (:366-375).  The last two run on a single GPU only. */
  int32_t algorithm;
  /* pathfinder (:366-375): num_paths = n.cores, max_lbfgs_iters = 200, single-path draws (CmdStan default 1000),
   * init = 0.01; num_elbo_draws = elbo_samples (50 for pathfinder, :111-112; <= 50), history_size = lbfgs_history (<= 25),
   * PSIS draws = n_draws.  mle_init does not apply (:109). */
  int32_t pf_num_paths;
  int32_t pf_max_lbfgs_iters;
  int32_t pf_single_draws;
  double pf_init_radius;
} bvg_config;

typedef enum { BVG_ALG_MEANFIELD = 0, BVG_ALG_FULLRANK = 1, BVG_ALG_PATHFINDER = 2 } bvg_algorithm;

typedef struct bvg_fit bvg_fit;

int bvg_version(void);
const char *bvg_last_error(void);
int bvg_device_count(void);
void bvg_config_default(bvg_config *cfg);

/* Basis builder (:265-288): phi_out = thin Q of column-pivoted Householder QR of
 * kernel(|x_s - c_i|^2, lscale).  coords M x 2 (already z-scored, :160), centers k x 2. */
int bvg_basis_build(int device, const double *coords, int64_t M, const double *centers, int k,
                    double lscale, int kernel, double nu, double period, double *phi_out);

/* Centroids (:200-204): stats::kmeans(spatial_mtx, centers = k, iter.max = 100, nstart = 10)$centers on the device.
 * Lloyd-Forgy from `nstart` sets of k distinct rows drawn with Philox (key = seed; the reference draws its starts from
 * R's ambient RNG, so its centroids are not reproducible and parity on them is statistical); the start with the
 * smallest total within-cluster sum of squares wins.  Sums are 64-bit fixed point: the result is bit-reproducible.
 * coords M x 2 col-major, already z-scored (:160; |x| < 256, M <= 2^22); centers_out k x 2 col-major;
 * tot_withinss and info = {winning start, its iterations, converged 0/1} may be NULL. */
int bvg_kmeans(int device, const double *coords, int64_t M, int k, int iter_max, int nstart, uint64_t seed,
               double *centers_out, double *tot_withinss, int32_t *info);

/* lscale.estimator = "variogram" (:208-260): per gene, gstat::variogram(z ~ 1) (cutoff = bounding-box diagonal / 3, 15 distance
 * classes) and gstat::fit.variogram of vgm(0.8 var, "Mat", median(dist), 0.2 var, kappa = kernel.smoothness) by weighted least
 * squares (weights N_h / h^2) [U]; *lscale_out = median of the fitted ranges, NaN ranges (failed fits, constant genes) dropped.
 * coords M x 2 col-major, scaled (:209-211); expr G x M col-major = R's expr_mtx[naive.hvgs, ] (:174); kappa in {0.5, 1.5, 2.5};
 * ranges_out (G) may be NULL. */
int bvg_variogram_lscale(int device, const double *coords, int64_t M, const double *expr, int64_t G, double kappa,
                         double *ranges_out, double *lscale_out);

/* The step after the fit: clusterSVGsBayes() (R/clusterSVGsBayes.R:63-319) = ADVI on inst/clusterSVGs.stan, a diagonal-covariance
 * Gaussian mixture over the PCA embedding X of the SVGs (N x D COLUMN-major, as R holds svg_mtx_pca$x), K = n.clust clusters.
 * Unconstrained vector (declaration order, :8-12): simplex[K] theta (K - 1 stick-breaking coordinates [U]), matrix[K, D] mu
 * (column-major), array[K] vector<lower=0>[D] sigma (log); P = K - 1 + 2 K D.
 * bvg_gmm_log_prob: the parity hook (log_prob / grad_log_prob of that program).
 * bvg_gmm_fit: mod$variational(init = 0, algorithm = "meanfield" | "fullrank", iter, draws, elbo_samples) (:153-160) followed by
 * the posterior pass of :201-291.  opts[7] = {n_iter (30000), n_draws (1000), elbo_samples (300), eval_elbo (100), adapt_iter (50),
 * eta (<= 0: adapt), tol_rel_obj (0.01)}; mu_out [P]; var_out [P] (omega) or [P (P + 1) / 2] (packed row-major Cholesky factor);
 * draws_out [n_draws][K + 2 K D] row-major (theta | mu column-major | sigma) or NULL; soft_out [N][K] row-major = the soft
 * cluster assignment probabilities (:215-239); stats_out[8] = {eta, iterations, converged, last ELBO, log-likelihood (:281-284),
 * seconds, gradient evaluations, ELBO estimates}.  One CTA runs every stretch of iterations without returning to the host. */
int bvg_gmm_log_prob(int device, const double *X, int64_t N, int D, int K, const double *theta, int jacobian, int propto,
                     double *lp, double *grad);
int bvg_gmm_fit(int device, const double *X, int64_t N, int D, int K, int fullrank, uint64_t seed, const double *opts,
                double *mu_out, double *var_out, double *draws_out, double *soft_out, double *stats_out);
/* algorithm = "pathfinder" of clusterSVGsBayes() (R/clusterSVGsBayes.R:162-170: mod$pathfinder(init = 0.01, draws, num_elbo_draws,
 * max_lbfgs_iters = 500, history_size = 100)): the host runs Pathfinder's L-BFGS paths and its (2 J)-dimensional algebra, as CmdStan
 * does; every log density / gradient is evaluated here with the embedding resident on the device.  bvg_gmm_open keeps X (N x D
 * column-major) on the device; bvg_gmm_eval evaluates log_prob<propto, jacobian> at B unconstrained points in ONE launch (one CTA
 * per point: the num_elbo_draws draws of an iterate), gradients ([B][P]) optional; bvg_gmm_posterior is the posterior pass of
 * :201-291 for GIVEN unconstrained draws [nd][P] (the PSIS-resampled draws): constrained draws [nd][K + 2 K D] or NULL, soft
 * assignments [N][K] row-major, log-likelihood. */
int bvg_gmm_open(int device, const double *X, int64_t N, int D, int K, void **handle);
void bvg_gmm_close(void *handle);
int bvg_gmm_eval(void *handle, const double *thetas, int B, int jacobian, int propto, double *lp_out, double *grad_out);
int bvg_gmm_posterior(void *handle, const double *thetas, int nd, double *draws_out, double *soft_out, double *loglik_out);

int bvg_fit_create(const bvg_config *cfg, bvg_fit **out);
void bvg_fit_destroy(bvg_fit *);
/* bvg_fit_destroy keeps the large device blocks of a fit (the staged expression matrix, the H2D staging buffers, the L-BFGS work
 * space: up to 16 blocks / 1.5 GB per process) for the next fit that asks for the same sizes -- allocating and freeing them costs
 * the driver up to 0.4 s, twice a whole fit.  This hands them back to the driver (e.g. before another library needs the memory). */
void bvg_release_cached_memory(void);

/* data_list$phi (M x k) and data_list$y (the gene-major long vector = M x G column-major).
 * spot_id / gene_id are implied (N = M*G). */
int bvg_fit_set_phi(bvg_fit *, const double *phi, int64_t M, int k);
int bvg_fit_set_y_f64(bvg_fit *, const double *y, int64_t M, int64_t G);
int bvg_fit_set_y_i32(bvg_fit *, const int32_t *y, int64_t M, int64_t G);
/* data_list$gene_depths (:303): log of each gene's total count, length G; only with cfg.depth_adjust = 1.
 * May be called before or after bvg_fit_set_y; must be set before any evaluation. */
int bvg_fit_set_gene_depths(bvg_fit *, const double *gene_depths, int64_t G);
int bvg_fit_dim(bvg_fit *, int64_t *D);
/* shape of the staged model: M spots, G (local) genes, k basis functions, D unconstrained parameters; any pointer may
 * be NULL.  Lets a binding size and validate its buffers from the handle instead of trusting caller-supplied sizes. */
int bvg_fit_shape(bvg_fit *, int64_t *M, int64_t *G, int *k, int64_t *D);

/* multi-GPU (one process per GPU of one node; the reference has no counterpart -- its fit is one
 * CPU thread, R/findSpatiallyVariableFeaturesBayes.R:123-129).  Two transports for the one exchange
 * step of the path, the sum of 2k+10 cross-gene scalars per evaluation; choose ONE per fit:
 *  - peer memory (default of the host wrapper): every rank calls bvg_fit_peer_handle (64 opaque bytes,
 *    a CUDA IPC handle of its exchange buffer), the host all-gathers them in rank order, every rank
 *    calls bvg_fit_peer_init.  The all-reduce then runs inside the per-gene kernel itself (stores into
 *    the peers' buffers over NVLink), with no extra launch.
 *  - NCCL: rank 0 calls bvg_comm_unique_id, the host broadcasts the bytes, every rank calls
 *    bvg_fit_comm_init.  NCCL is resolved with dlopen("libnccl.so.2") at that point. */
int bvg_fit_peer_handle(bvg_fit *, void *out, size_t cap, size_t *len);
int bvg_fit_peer_init(bvg_fit *, const void *handles /* world_size x 64 bytes, rank order */, size_t len);
int bvg_comm_unique_id(void *out, size_t cap, size_t *len);
int bvg_fit_comm_init(bvg_fit *, const void *unique_id, size_t len);

/* log density and gradient at an unconstrained vector in Stan declaration order (local layout
 * when sharded).  Stands in for cmdstanr's $log_prob / $grad_log_prob.  grad may be NULL. */
int bvg_log_prob(bvg_fit *, const double *theta, int jacobian, int propto, double *lp, double *grad);

/* whole path: [L-BFGS init] -> eta adaptation -> stochastic gradient ascent -> draws -> summary */
int bvg_fit_run(bvg_fit *);

/* the stages of bvg_fit_run, individually callable */
int bvg_fit_mle(bvg_fit *, const double *theta0 /* NULL = 0 */, double *theta_out /* may be NULL */);
int bvg_fit_advi(bvg_fit *, const double *theta_init /* NULL = MLE result or 0 */);
int bvg_fit_summarize(bvg_fit *);

/* finer-grained hooks used by the parity tests and the benchmark */
int bvg_fit_set_variational(bvg_fit *, const double *mu, const double *omega /* NULL = 0 */);
int bvg_fit_get_variational(bvg_fit *, double *mu, double *omega);
/* algorithm = fullrank: the Cholesky factor of the approximation, PACKED lower-triangular row-major (row i starts at
 * i (i + 1) / 2; D (D + 1) / 2 doubles).  bvg_fit_set_variational installs (mu, L = I) [U: normal_fullrank's start];
 * bvg_fit_get_variational reports omega = log |L_ii|. */
/* algorithm = pathfinder: ONE path (the parity hook next to the CPU restatement): amp [pf_single_draws][G] row-major
 * amplitude draws of the path's best approximation, lp_ratio [pf_single_draws] = log p - log q of each draw,
 * info[5] = {L-BFGS iterations, L-BFGS termination code, best iterate, curvature pairs in it, its ELBO estimate},
 * elbos (may be NULL) [pf_max_lbfgs_iters + 1] the ELBO estimate at every iterate (NaN where none was made).
 * bvg_fit_run runs pf_num_paths of them, Pareto-smoothed-importance resamples n_draws draws and summarises those;
 * afterwards bvg_fit_get_trace rows are (path, best iterate, best ELBO, L-BFGS iterations) and stats.eta = Pareto k. */
int bvg_fit_pathfinder_single(bvg_fit *, int path, double *amp, double *lp_ratio, double *info, double *elbos);
int bvg_fit_get_cholesky(bvg_fit *, double *L_packed);
int bvg_fit_set_cholesky(bvg_fit *, const double *L_packed);
/* n stochastic-gradient steps (iteration counter continues from the last reset);
 * ms_out (optional) = device time from CUDA events on the engine's stream */
int bvg_fit_advi_steps(bvg_fit *, int n, double eta, int reset_counters, float *ms_out);
int bvg_fit_elbo(bvg_fit *, int n_samples, double *elbo_out, float *ms_out);
/* time n launches of the likelihood kernel alone; flush_l2 = 1 evicts L2 before each launch (event pair around each), 2 (fused
 * tcgen05 engine) launches back to back over four rotating copies of Y -- an input larger than L2 -- with one event pair */
int bvg_fit_time_lik_kernel(bvg_fit *, int n, int flush_l2, int backward, float *ms_avg_out);

/* evict the L2 (writes a 512 MB scratch buffer) -- benchmark hygiene between timed steps */
int bvg_fit_flush_l2(bvg_fit *);

/* results */
int bvg_fit_get_summary(bvg_fit *, double *out /* G x 8 col-major: mean, median, sd, mad, q5, q95, dispersion, rank */);
int bvg_fit_get_draws(bvg_fit *, double *out /* n_draws x G col-major amplitude draws */);
/* save.model = TRUE (:483-489, R/extractModel.R:21-25): the content of the CmdStanVB object's draws, i.e. of CmdStan's
 * variational output CSV.  out is n x ncol COLUMN-major with columns lp__ (0), log_p__, log_g__, the D parameters
 * constrained in declaration order, then the transformed parameters beta0[G] (hierarchical intercept only) and
 * amplitude_sq[G].  Draw d uses Philox stream 3, counter d, so its amplitude columns equal bvg_fit_get_draws. */
int bvg_fit_posterior_ncol(bvg_fit *, int64_t *ncol);
int bvg_fit_get_posterior(bvg_fit *, int n, double *out);
int bvg_fit_get_trace(bvg_fit *, double *out /* cap x 4 row-major: iter, elbo, delta_mean, delta_med */, int cap, int *n);

typedef struct {
  double eta;
  int32_t advi_iters, grad_evals, elbo_evals, converged; /* converged: 1 mean, 2 median, 3 both, 0 max-iter */
  int32_t mle_iters, mle_fevals, mle_term;
  double mle_lp, elbo;
  double sec_h2d, sec_mle, sec_adapt, sec_sga, sec_draws, sec_total;
  int64_t kernel_launches;
  int32_t engine_used; /* bvg_engine actually running the likelihood */
  int32_t mle_fp64_from; /* fast precision: L-BFGS iteration from which the fp64 evaluator took over (0 = never) */
} bvg_fit_stats;
int bvg_fit_get_stats(bvg_fit *, bvg_fit_stats *out);

/* diagnostics (not needed by a binding; used by scripts/ and profiles/) */
int bvg_debug_stage_times(bvg_fit *, int n, float *out4 /* ms: prep, likelihood, gene+finish, - */);
int bvg_debug_gene_trace(bvg_fit *, long long *out16 /* %globaltimer ns stamps inside k_gene */, int ahead);
int bvg_debug_time_gene(bvg_fit *, int n, int mode /* 0 grad out, 1 advi update, 3 lp only */, int fuse /* 0 | 3 */, int ahead, float *ms_avg);
int bvg_debug_tc_trace(bvg_fit *, long long *out /* [chunk][8] clock64 stamps of CTA (0,0) */, int nchunk_cap);
int bvg_debug_persist_trace(bvg_fit *, int n_iter, long long *out /* [iteration][8] clock64 stamps of CTA 0 of k_advi_persist */, int cap_iters);

#ifdef __cplusplus
}
#endif
#endif

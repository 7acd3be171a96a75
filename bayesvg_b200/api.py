"""Host-side mirror of the reference's R interface for the SVG hot path.

`findSpatiallyVariableFeaturesBayes` keeps the argument names (dots -> underscores), defaults,
validation messages and output table of /root/reference/R/findSpatiallyVariableFeaturesBayes.R:78-508;
the numeric middle (basis QR, L-BFGS init, mean-field ADVI, draws, summaries) is one sequence of
C-ABI calls into libbayesvg_b200.so instead of cmdstanr/CmdStan.  `SpatialData` is a minimal stand-in
for the Seurat / SpatialExperiment containers (R is not available in this image; the R-side shim is
in r_shim/ and INTEGRATION.md).
"""
import time

import numpy as np
import pandas as pd

from . import engine as E

_SUMMARY_COLS = E.SUMMARY_COLUMNS


class SpatialData:
    """coords: M x 2; data / counts: G x M (genes x spots) like Seurat's assay layers."""

    def __init__(self, coords, data=None, counts=None, genes=None, spots=None):
        self.coords = np.asarray(coords, dtype=np.float64)
        self.data = None if data is None else np.asarray(data)
        self.counts = None if counts is None else np.asarray(counts)
        ref = self.data if self.data is not None else self.counts
        if ref is None:
            raise ValueError("provide at least one of data / counts")
        G, M = ref.shape
        if self.coords.shape[0] != M:
            raise ValueError("coords rows must equal the number of spots")
        self.genes = [f"gene{g + 1}" for g in range(G)] if genes is None else list(genes)
        self.spots = [f"spot{s + 1}" for s in range(M)] if spots is None else list(spots)
        self.meta_features = pd.DataFrame(index=pd.Index(self.genes, name="gene"))
        self.variable_features = []
        self.misc = {}

    def rownames(self):
        return self.genes


class VariationalFit:
    """What `save.model = TRUE` keeps (:483-489) and `extractModel()` returns: the stand-in for cmdstanr's CmdStanVB
    object.  It holds the n.draws x ncol draws matrix of the whole model (the content of CmdStan's variational output
    CSV: lp__, log_p__, log_g__, parameters constrained in declaration order, transformed parameters) plus the
    variational mean / log-sd, the ELBO trace and the run statistics, and mirrors the CmdStanVB methods the
    reference and its users call: `$draws(variables)`, `$summary(variables)`, `$lp()`, `$lp_approx()`,
    `$num_draws()`, `$save_output_files()` (here `save_csv`)."""

    def __init__(self, columns, draws, mu=None, omega=None, trace=None, stats=None, **extra):
        draws = np.asarray(draws)
        if draws.ndim != 2 or draws.shape[1] != len(columns):
            raise ValueError("draws must be n.draws x length(columns)")
        self.columns = list(columns)
        self._draws = draws
        self._index = {c: i for i, c in enumerate(self.columns)}
        self.mu, self.omega, self.trace, self.stats = mu, omega, trace, stats
        self.extra = extra

    def _select(self, variables):
        if variables is None:
            return list(range(len(self.columns)))
        if isinstance(variables, str):
            variables = [variables]
        idx = []
        for v in variables:
            if v in self._index:               # a scalar or one element: "sigma_y", "amplitude[3]"
                idx.append(self._index[v])
                continue
            hit = [i for i, c in enumerate(self.columns) if c.startswith(v + "[")]   # a whole container: "amplitude"
            if not hit:
                raise KeyError(f"Can't find the following variable(s) in the output: {v}")
            idx.extend(hit)
        return idx

    def num_draws(self):
        return self._draws.shape[0]

    def mean_row(self):
        """The first data row of CmdStan's variational CSV: the mean of the approximation, constrained, with
        lp__ = log_p__ = log_g__ = 0 [U] -- what cmdstanr drops before building `$draws()`."""
        if self.mu is None:
            return None
        D = len(self.mu)
        names = self.columns[3:3 + D]
        pos = np.array([n.startswith(("sigma_", "amplitude[")) or n == "phi_nb" for n in names])
        th = np.where(pos, np.exp(self.mu), self.mu)
        val = dict(zip(names, th))
        tp = []
        for c in self.columns[3 + D:]:
            g = c[c.index("[") + 1:-1]
            tp.append(val["mu_beta0"] + val["sigma_beta0"] * val[f"z_beta0[{g}]"] if c.startswith("beta0[")
                      else val[f"amplitude[{g}]"] ** 2)
        return np.concatenate([np.zeros(3), th, np.array(tp)])

    def draws(self, variables=None):
        """`fit$draws(variables, format = "draws_df")`: a DataFrame, one row per draw."""
        idx = self._select(variables)
        return pd.DataFrame(self._draws[:, idx], columns=[self.columns[i] for i in idx])

    def lp(self):
        """`fit$lp()`: log_p__, the log density of the model at each draw (jacobian = TRUE, all constants)."""
        return self._draws[:, self._index["log_p__"]].copy()

    def lp_approx(self):
        """`fit$lp_approx()`: log_g__, the log density of the approximation at each draw up to a constant."""
        return self._draws[:, self._index["log_g__"]].copy()

    def summary(self, variables=None):
        """`fit$summary(variables)` for a variational fit = posterior::default_summary_measures() [U]:
        mean, median, sd, mad (constant 1.4826), q5, q95 with R's type-7 quantiles (:419)."""
        idx = self._select(variables)
        x = self._draws[:, idx]
        med = np.median(x, axis=0)
        out = pd.DataFrame({
            "variable": [self.columns[i] for i in idx],
            "mean": x.mean(0), "median": med, "sd": x.std(0, ddof=1) if x.shape[0] > 1 else np.full(len(idx), np.nan),
            "mad": 1.4826 * np.median(np.abs(x - med), axis=0),
            "q5": np.quantile(x, 0.05, axis=0), "q95": np.quantile(x, 0.95, axis=0)})
        return out

    def save_csv(self, path):
        """CmdStan's variational output CSV: the header, then the mean of the approximation as the first row
        (lp__ = log_p__ = log_g__ = 0) [U], then the draws."""
        with open(path, "w") as fh:
            fh.write("# method = variational (bayesvg_b200)\n")
            # CmdStan writes container elements as name.i.j; cmdstanr turns them into name[i,j] when it reads the file
            fh.write(",".join(c.replace("[", ".").replace(",", ".").replace("]", "") for c in self.columns) + "\n")
            if self.mu is not None:
                fh.write(",".join(repr(float(v)) for v in self.mean_row()) + "\n")
            for row in self._draws:
                fh.write(",".join(repr(float(v)) for v in row) + "\n")
        return path


def extractModel(obj=None):
    """R/extractModel.R:18-27."""
    if obj is None:
        _abort("Argument obj must be non-NULL.")
    return obj.misc.get("model_fit")


def scale_coords(coords):
    """coop::scaler: column z-score, sd with n-1 (:160)."""
    X = np.asarray(coords, dtype=np.float64)[:, :2]
    return (X - X.mean(0)) / X.std(0, ddof=1)


def kmeans_centers(X, k, iter_max=100, nstart=10, rng=None, device=0):
    """stats::kmeans(X, centers = k, iter.max = 100, nstart = 10)$centers (:200-204), on the device (bvg_kmeans).

    R runs Hartigan-Wong from `nstart` random starts drawn from the ambient RNG and keeps the best; here: Lloyd
    iterations from `nstart` Philox-drawn starts, best within-SS kept.  The Philox key comes from `rng` (a numpy
    Generator or an int); `rng=None` takes it from numpy's global state, mirroring the reference's un-seeded call."""
    X = np.asarray(X, dtype=np.float64)
    if k > X.shape[0]:
        raise ValueError("more cluster centers than distinct data points.")
    if rng is None:
        seed = int(np.random.randint(0, 2**31 - 1))
    elif isinstance(rng, (int, np.integer)):
        seed = int(rng)
    else:
        seed = int(rng.integers(0, 2**63 - 1))
    return E.kmeans(X, k, iter_max, nstart, seed, device)


def length_scale_kmeans(centers):
    """median(dist(centers)) over the k(k-1)/2 pairs (:205-207)."""
    C = np.asarray(centers)
    d = np.sqrt(((C[:, None, :] - C[None]) ** 2).sum(-1))
    return float(np.median(d[np.triu_indices(C.shape[0], 1)]))


def _abort(msg):
    raise ValueError(msg)


def findSpatiallyVariableFeaturesBayes(sp_obj=None, naive_hvgs=None, likelihood="gaussian", lscale_estimator="kmeans",
                                       n_iter=3000, kernel="exp_quad", kernel_smoothness=1.5, kernel_period=100,
                                       n_basis_fns=20, algorithm="meanfield", mle_init=True, mle_iter=1000,
                                       gene_depth_adjust=False, n_draws=1000, elbo_samples=None, opencl_params=None,
                                       n_cores=2, random_seed=312, verbose=True, save_model=False, *,
                                       centers=None, precision="fast", engine="auto", device=0, kmeans_rng=None,
                                       shard=None):
    """Drop-in for the reference function (:78-97).  Extra keyword-only arguments: `centers` (k x 2
    precomputed centroids in scaled coordinates, so that Phi is reproducible), `precision`
    ("fast" | "fp64"), `engine`, `device`, `shard` (a bayesvg_b200.dist.Shard for multi-GPU)."""
    # check & parse inputs (:99-138)
    if sp_obj is None:
        _abort("Please provide a spatial data object to findSpatiallyVariableFeaturesBayes().")
    if not isinstance(sp_obj, SpatialData):
        _abort("Please provide an object of class Seurat or SpatialExperiment.")
    if naive_hvgs is None:
        _abort("Please identify a set of naive HVGs prior to running findSpatiallyVariableFeaturesBayes().")
    likelihood = likelihood.lower()
    if likelihood not in ("gaussian", "nb"):
        _abort("Please provide a valid likelihood.")
    kernel = kernel.lower()
    if kernel not in ("exp_quad", "matern", "periodic"):
        _abort("Please provide a valid covariance kernel.")
    if kernel == "matern" and kernel_smoothness not in (0.5, 1.5, 2.5):
        _abort("When utilizing the Matern kernel you must provide a valid smoothness parameter value.")
    algorithm = algorithm.lower()
    if algorithm not in ("meanfield", "fullrank", "pathfinder"):
        _abort("Please provide a valid variational inference approximation algorithm.")
    if elbo_samples is None:
        elbo_samples = 50 if algorithm == "pathfinder" else 150
    # accepted by the reference, rejected clearly here (north_star: no fallback)
    if mle_init and algorithm == "pathfinder" and verbose:   # :109
        print("! Initialization at the regularized MLE is not supported when using the Pathfinder algorithm.")
    if algorithm != "meanfield" and shard is not None:
        raise NotImplementedError(f"algorithm = '{algorithm}' is not built for gene-sharded fits; use 'meanfield'.")
    if algorithm == "pathfinder" and save_model:
        raise NotImplementedError("save.model = TRUE is not built for algorithm = 'pathfinder'.")
    if shard is not None and device == 0:
        # one process per GPU (torchrun): the default device of a sharded fit is this rank's GPU, not GPU 0 for everyone
        import os
        device = int(os.environ.get("LOCAL_RANK", shard.rank))
    if opencl_params is not None:
        raise NotImplementedError("opencl.params is meaningless here: the fit always runs on the CUDA device.")
    if lscale_estimator not in ("kmeans", "variogram"):
        _abort("Please provide a valid length-scale estimator.")
    if lscale_estimator == "variogram" and kernel_smoothness not in (0.5, 1.5, 2.5):
        raise NotImplementedError("lscale.estimator = 'variogram' is built for kernel.smoothness in {0.5, 1.5, 2.5} only.")
    t_start = time.time()
    # coordinates (:154-161)
    X = scale_coords(sp_obj.coords)
    # expression (:163-193)
    layer = sp_obj.data if likelihood == "gaussian" else sp_obj.counts
    if layer is None:
        _abort("the object has no '%s' layer" % ("data" if likelihood == "gaussian" else "counts"))
    gidx = {g: i for i, g in enumerate(sp_obj.genes)}
    missing = [g for g in naive_hvgs if g not in gidx]
    if missing:
        raise KeyError(f"subscript out of bounds: {missing[:3]}")
    rows = [gidx[g] for g in naive_hvgs]
    expr = np.asarray(layer)[rows, :]  # G x M
    if likelihood == "gaussian":
        v = expr.astype(np.float64)
        Y = np.asfortranarray(((v - v.mean()) / v.std(ddof=1)).T)  # global z-score (:185); M x G
    else:
        Y = np.asfortranarray(expr.astype(np.int32).T)
    # gene.depth.adjust (:293-303): gene_depths = log(rowSums(counts[naive.hvgs, ])) -> approxGP2 / approxGP3
    gene_depths = None
    if gene_depth_adjust:
        if sp_obj.counts is None:
            _abort("gene.depth.adjust = TRUE needs the 'counts' layer")
        with np.errstate(divide="ignore"):
            gene_depths = np.log(np.asarray(sp_obj.counts)[rows, :].sum(axis=1).astype(np.float64))
    # length-scale + basis (:200-288)
    if centers is None:
        centers = kmeans_centers(X, n_basis_fns, 100, 10, kmeans_rng, device)
    centers = np.asarray(centers, dtype=np.float64)
    if centers.shape != (n_basis_fns, 2):
        _abort("centers must be n.basis.fns x 2")
    if lscale_estimator == "kmeans":
        lscale = length_scale_kmeans(centers)
    else:   # :208-260: median fitted Matern range of the per-gene empirical variograms (expr_mtx rows, not the z-scored vector)
        lscale = E.variogram_lscale(X, np.asarray(expr, dtype=np.float64), kernel_smoothness, device)
    if verbose:
        print(f"i Estimated length-scale: {round(lscale, 4)}")
    phi = E.basis_build(X, centers, lscale, kernel, kernel_smoothness, float(kernel_period), device)
    # fit (:337-413)
    cfg = dict(n_iter=int(n_iter), mle_init=int(bool(mle_init)), mle_iter=int(mle_iter), n_draws=int(n_draws),
               elbo_samples=int(elbo_samples), seed=int(random_seed), verbose=int(bool(verbose)), device=device)
    if algorithm == "pathfinder":   # :366-375: num_paths = n.cores, max_lbfgs_iters = 200, history_size = 25, init = 0.01
        cfg.update(pf_num_paths=int(n_cores), pf_max_lbfgs_iters=200, pf_init_radius=0.01, lbfgs_history=25)
    if shard is not None:
        from . import dist
        table, model = dist.fit_sharded(phi, Y, shard, likelihood, precision, engine, cfg, gene_depths=gene_depths,
                                        save_model=bool(save_model), n_draws=int(n_draws), genes=list(naive_hvgs),
                                        extra=dict(phi=phi, lscale=lscale, centers=centers))
    else:
        fit = E.Fit(phi, Y, likelihood, precision, engine, gene_depths=gene_depths, algorithm=algorithm, **cfg)
        table = fit.run()
        model = None
        if save_model:
            mu, om = fit.get_variational()
            model = VariationalFit(fit.posterior_columns(), fit.posterior(int(n_draws)), mu=mu, omega=om, trace=fit.trace(),
                                   stats=fit.stats(), phi=phi, lscale=lscale, centers=centers, genes=list(naive_hvgs))
        stats = fit.stats()
        fit.close()
        if verbose:
            if algorithm == "pathfinder":
                print(f"v Pathfinder: {n_cores} paths, {stats['advi_iters']} L-BFGS iterations, best ELBO = {stats['elbo']:.6g}, Pareto k = {stats['eta']:.3g}")
            else:
                print(f"v ADVI ({algorithm}): eta = {stats['eta']}, {stats['advi_iters']} iterations, ELBO = {stats['elbo']:.6g}")
    # summarise (:419-432)
    summ = pd.DataFrame(table[:, :7], columns=_SUMMARY_COLS[:7])
    summ.insert(0, "gene", list(naive_hvgs))
    summ = summ.sort_values("amplitude_mean", ascending=False, kind="stable").reset_index(drop=True)
    summ["amplitude_mean_rank"] = np.arange(1, len(summ) + 1)
    summ.index = pd.Index(summ["gene"], name=None)
    if verbose:
        print("v Posterior summarization complete.")
    # write-back (:436-481): untested genes get NA rows, order = rownames(sp.obj)
    full = summ.reindex(sp_obj.genes)
    full["gene"] = sp_obj.genes
    for c in _SUMMARY_COLS:
        sp_obj.meta_features[c] = full[c].to_numpy()
    if save_model:
        sp_obj.misc["model_fit"] = model
    if verbose:
        print(f"v bayesVG modeling of {len(naive_hvgs)} naive HVGs completed in {round(time.time() - t_start, 3)} seconds.")
    return sp_obj


def getBayesianGeneStats(obj, sort_values=True):
    """R/getBayesianGeneStats.R:60-68: drop untested genes, sort by amplitude_mean descending."""
    df = obj.meta_features.copy()
    df.insert(0, "gene", df.index)
    df = df[~df["amplitude_mean"].isna()]
    if sort_values:
        df = df.sort_values("amplitude_mean", ascending=False, kind="stable")
    return df


def classifySVGs(obj, selection_method="rank", n_SVG=1000, quantile_SVG=0.75, cutoff=0.1):
    """R/classifySVGs.R:60-82: label the top genes by amplitude_mean_rank / quantile / cutoff."""
    df = obj.meta_features
    if "amplitude_mean_rank" not in df:
        _abort("Please run findSpatiallyVariableFeaturesBayes() first.")
    if selection_method == "rank":
        sel = df["amplitude_mean_rank"] <= n_SVG
    elif selection_method == "quantile":
        sel = df["amplitude_mean"] >= df["amplitude_mean"].quantile(quantile_SVG)
    elif selection_method == "cutoff":
        sel = df["amplitude_mean"] >= cutoff
    else:
        _abort("Please provide a valid SVG selection method.")
    obj.meta_features["svg_status"] = np.where(df["amplitude_mean"].isna(), None, sel.fillna(False))
    order = df.loc[sel.fillna(False)].sort_values("amplitude_mean_rank").index
    obj.variable_features = list(order)
    return obj


def clusterSVGsBayes(sp_obj=None, svgs=None, n_clust=5, n_PCs=30, algorithm="fullrank", n_iter=30000, n_draws=1000,
                     elbo_samples=None, opencl_params=None, n_cores=2, random_seed=312, verbose=True, *, device=0):
    """Drop-in for R/clusterSVGsBayes.R:63-319: scale the SVGs' expression, PCA, fit the Bayesian Gaussian mixture of
    inst/clusterSVGs.stan by ADVI on the GPU, soft / hard cluster assignments, log-likelihood and BIC."""
    if sp_obj is None or svgs is None:
        _abort("All arguments to clusterSVGsBayes() must be supplied.")
    if not isinstance(sp_obj, SpatialData):
        _abort("Please provide an object of class Seurat or SpatialExperiment.")
    if n_clust <= 1:
        _abort("Please provide a valid number of clusters.")
    algorithm = algorithm.lower()
    if algorithm not in ("meanfield", "fullrank", "pathfinder"):
        _abort("Please provide a valid variational inference approximation algorithm.")
    if opencl_params is not None:
        raise NotImplementedError("opencl.params is meaningless here: the fit always runs on the CUDA device.")
    if algorithm == "meanfield" and verbose:
        print("! The meanfield VI algorithm often does not perform well on clustering problems do to multi-modality and high "
              "correlation of the posterior. Use caution and consider utilizing the default fullrank VI algorithm instead.")
    if elbo_samples is None:
        elbo_samples = 100 if algorithm == "pathfinder" else 300                                 # :19-25
    t_start = time.time()
    gidx = {g: i for i, g in enumerate(sp_obj.genes)}
    rows = [gidx[g] for g in svgs]
    expr = np.asarray(sp_obj.data, dtype=np.float64)[rows, :]                                  # :125-132 layer "data", svgs x spots
    expr = (expr - expr.mean(1, keepdims=True)) / expr.std(1, ddof=1, keepdims=True)           # t(coop::scaler(t(expr_mtx)))
    # irlba::prcomp_irlba(expr_mtx, n = n.PCs, center = FALSE, scale. = FALSE)$x = U S of the leading singular triplets (:134-137)
    U, S, _ = np.linalg.svd(expr, full_matrices=False)
    n_pc = min(int(n_PCs), len(S))
    emb = U[:, :n_pc] * S[:n_pc]
    if algorithm == "pathfinder":
        # mod$pathfinder(init = 0.01, draws = n.draws, num_elbo_draws = elbo.samples, max_lbfgs_iters = 500, history_size = 100) (:162-170)
        from .gmm_pathfinder import gmm_pathfinder
        res = gmm_pathfinder(emb, n_clust, n_draws=n_draws, elbo_samples=elbo_samples, seed=random_seed, num_paths=4,
                             max_lbfgs_iters=500, history=100, init_radius=0.01, device=device)
    else:
        res = E.gmm_fit(emb, n_clust, algorithm, n_iter, n_draws, elbo_samples, random_seed, device=device)
    soft = res["soft"]
    K, D, N = int(n_clust), n_pc, emb.shape[0]
    cluster_df = pd.DataFrame(soft, columns=[f"prob_cluster_{j + 1}" for j in range(K)])
    cluster_df.insert(0, "gene", list(svgs))
    cluster_df["assigned_cluster"] = soft.argmax(1) + 1                                          # :262-266 which.max
    ll = res["info"]["log_likelihood"]                                                           # :281-284
    p = (K - 1) + 2 * K * D                                                                      # :286
    cols = ([f"theta[{j + 1}]" for j in range(K)] + [f"mu[{j + 1},{d + 1}]" for d in range(D) for j in range(K)]
            + [f"sigma[{j + 1},{d + 1}]" for j in range(K) for d in range(D)])
    model_fit = VariationalFit(cols, res["draws"], trace=None, stats=res["info"], mu_unconstrained=res["mu"], var=res["var"])
    if verbose:
        i = res["info"]
        if algorithm == "pathfinder":
            print(f"v Pathfinder: {len(i['paths'])} paths, {i['iters']} L-BFGS iterations, best ELBO = {i['elbo']:.6g}, Pareto k = {i['khat']:.2f}")
        else:
            print(f"v ADVI ({algorithm}): eta = {i['eta']}, {i['iters']} iterations, ELBO = {i['elbo']:.6g}")
        print(f"v bayesVG clustering of {len(svgs)} SVGs completed in {round(time.time() - t_start, 3)} seconds.")
    return dict(pca_embedding=emb, cluster_df=cluster_df, model_fit=model_fit, log_likelihood=ll, BIC=-2 * ll + p * np.log(N))

"""Thin object wrapper over the C ABI: one `Fit` = one bvg_fit handle (one gene shard on one GPU)."""
import ctypes as C

import numpy as np

from . import _lib as B

SUMMARY_COLUMNS = ["amplitude_mean", "amplitude_median", "amplitude_sd", "amplitude_mad", "amplitude_ci_ll",
                   "amplitude_ci_ul", "amplitude_dispersion", "amplitude_mean_rank"]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_config(**kw):
    cfg = B.Config()
    B.lib().bvg_config_default(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown config field {k}")
        setattr(cfg, k, v)
    return cfg


def basis_build(coords, centers, lscale, kernel="exp_quad", nu=1.5, period=100.0, device=0):
    """Phi = qr.Q(qr(kernel(|x - c|^2))) on the GPU (findSpatiallyVariableFeaturesBayes.R:265-288)."""
    X = np.asfortranarray(coords, dtype=np.float64)
    Cc = np.asfortranarray(centers, dtype=np.float64)
    if X.ndim != 2 or X.shape[1] != 2 or Cc.ndim != 2 or Cc.shape[1] != 2:
        raise ValueError("coords must be M x 2 and centers k x 2")
    if kernel not in B.KERNELS:
        raise ValueError("Please provide a valid covariance kernel.")
    M, k = X.shape[0], Cc.shape[0]
    out = np.empty((M, k), order="F")
    B.check(B.lib().bvg_basis_build(device, _dp(X), M, _dp(Cc), k, float(lscale), B.KERNELS[kernel], float(nu),
                                    float(period), _dp(out)))
    return out


def kmeans(coords, k, iter_max=100, nstart=10, seed=312, device=0, return_info=False):
    """stats::kmeans(coords, centers = k, iter.max, nstart)$centers on the GPU (findSpatiallyVariableFeaturesBayes.R:200-204)."""
    X = np.asfortranarray(coords, dtype=np.float64)
    if X.ndim != 2 or X.shape[1] != 2:
        raise ValueError("coords must be M x 2")
    out = np.empty((int(k), 2), order="F")
    ss = C.c_double()
    info = (C.c_int32 * 3)()
    B.check(B.lib().bvg_kmeans(device, _dp(X), X.shape[0], int(k), int(iter_max), int(nstart), int(seed), _dp(out),
                               C.byref(ss), info))
    if return_info:
        return out, dict(tot_withinss=ss.value, start=info[0], iters=info[1], converged=bool(info[2]))
    return out


def variogram_lscale(coords, expr, kappa=1.5, device=0, return_ranges=False):
    """lscale.estimator = "variogram" (findSpatiallyVariableFeaturesBayes.R:208-260) on the GPU: median over genes of the range of
    a Matern variogram model fitted to each gene's empirical semivariogram.  coords M x 2 (scaled), expr G x M (genes x spots)."""
    X = np.asfortranarray(coords, dtype=np.float64)
    Z = np.asfortranarray(expr, dtype=np.float64)
    if X.ndim != 2 or X.shape[1] != 2 or Z.ndim != 2 or Z.shape[1] != X.shape[0]:
        raise ValueError("coords must be M x 2 and expr G x M")
    ranges = np.empty(Z.shape[0])
    ls = C.c_double()
    B.check(B.lib().bvg_variogram_lscale(device, _dp(X), X.shape[0], _dp(Z), Z.shape[0], float(kappa), _dp(ranges), C.byref(ls)))
    return (ls.value, ranges) if return_ranges else ls.value


def gmm_log_prob(X, K, theta, jacobian=True, propto=False, grad=True, device=0):
    """log density (and gradient) of inst/clusterSVGs.stan at an unconstrained vector; X: N x D PCA embedding"""
    Xf = np.asfortranarray(X, dtype=np.float64)
    N, D = Xf.shape
    th = np.ascontiguousarray(theta, dtype=np.float64)
    if th.shape != (K - 1 + 2 * K * D,):
        raise ValueError("theta must have K - 1 + 2 K D entries")
    lp = C.c_double()
    g = np.empty_like(th) if grad else None
    B.check(B.lib().bvg_gmm_log_prob(device, _dp(Xf), N, D, int(K), _dp(th), int(jacobian), int(propto), C.byref(lp), _dp(g) if grad else None))
    return (lp.value, g) if grad else lp.value


def gmm_fit(X, K, algorithm="fullrank", n_iter=30000, n_draws=1000, elbo_samples=300, seed=312, eval_elbo=100, adapt_iter=50, eta=0.0,
            tol_rel_obj=0.01, device=0, return_draws=True):
    """mod$variational(...) on inst/clusterSVGs.stan + the posterior pass (R/clusterSVGsBayes.R:153-291) on the GPU"""
    if algorithm not in ("meanfield", "fullrank"):
        raise ValueError("Please provide a valid variational inference approximation algorithm.")
    Xf = np.asfortranarray(X, dtype=np.float64)
    N, D = Xf.shape
    K = int(K)
    P = K - 1 + 2 * K * D
    fr = algorithm == "fullrank"
    opts = np.array([n_iter, n_draws, elbo_samples, eval_elbo, adapt_iter, eta, tol_rel_obj], dtype=np.float64)
    mu, var = np.empty(P), np.empty(P * (P + 1) // 2 if fr else P)
    draws = np.empty((n_draws, K + 2 * K * D)) if return_draws else None
    soft, stats = np.empty((N, K)), np.empty(8)
    B.check(B.lib().bvg_gmm_fit(device, _dp(Xf), N, D, K, int(fr), int(seed), _dp(opts), _dp(mu), _dp(var),
                                _dp(draws) if return_draws else None, _dp(soft), _dp(stats)))
    info = dict(eta=stats[0], iters=int(stats[1]), converged=int(stats[2]), elbo=stats[3], log_likelihood=stats[4], seconds=stats[5],
                grad_evals=int(stats[6]), elbo_evals=int(stats[7]))
    return dict(mu=mu, var=var, draws=draws, soft=soft, info=info)


def posterior_column_names(nb, depth, G, k):
    """header of CmdStan's variational output CSV for approxGP (gaussian) / approxGP4 (nb) / approxGP2-3 (depth)"""
    vec = lambda name, n: [f"{name}[{i + 1}]" for i in range(n)]  # noqa: E731
    z = [f"z_alpha_t[{j + 1},{g + 1}]" for g in range(G) for j in range(k)]  # column-major
    hyper = ["mu_amplitude", "sigma_amplitude"]
    alpha = vec("mu_alpha", k) + vec("sigma_alpha", k) + z
    tail = "phi_nb" if nb else "sigma_y"
    if depth:      # approxGP2.stan:13-23 / approxGP3.stan:13-23
        cols = ["beta0", "beta1", tail] + vec("amplitude", G) + hyper + alpha
    elif nb:       # approxGP4.stan:12-23
        cols = ["mu_beta0", "sigma_beta0"] + vec("z_beta0", G) + [tail] + vec("amplitude", G) + hyper + alpha
    else:          # approxGP.stan:12-23
        cols = ["mu_beta0", "sigma_beta0"] + vec("z_beta0", G) + hyper + vec("amplitude", G) + alpha + [tail]
    tp = ([] if depth else vec("beta0", G)) + vec("amplitude_sq", G)
    return ["lp__", "log_p__", "log_g__"] + cols + tp


class Fit:
    """gene_depths (length G, log of each gene's total count) selects the depth-adjusted programs
    inst/approxGP2.stan / approxGP3.stan (gene.depth.adjust = TRUE in the reference)."""

    def __init__(self, phi, y, likelihood="gaussian", precision="fast", engine="auto", gene_depths=None,
                 algorithm="meanfield", **cfg):
        lik = {"gaussian": B.GAUSSIAN, "nb": B.NB}.get(likelihood)
        if lik is None:
            raise ValueError("Please provide a valid likelihood.")
        alg = {"meanfield": B.ALG_MEANFIELD, "fullrank": B.ALG_FULLRANK, "pathfinder": B.ALG_PATHFINDER}.get(algorithm)
        if alg is None:
            raise ValueError("Please provide a valid variational inference approximation algorithm.")
        self.cfg = default_config(algorithm=alg, likelihood=lik, precision={"fp64": B.PREC_FP64, "fast": B.PREC_FAST}[precision],
                                  engine={"auto": B.ENGINE_AUTO, "simt": B.ENGINE_SIMT, "tcgen05": B.ENGINE_TCGEN05}[engine],
                                  depth_adjust=int(gene_depths is not None), **cfg)
        self._h = C.c_void_p()
        B.check(B.lib().bvg_fit_create(C.byref(self.cfg), C.byref(self._h)))
        phi = np.asfortranarray(phi, dtype=np.float64)
        self.M, self.k = phi.shape
        B.check(B.lib().bvg_fit_set_phi(self._h, _dp(phi), self.M, self.k))
        if gene_depths is not None:
            self.set_gene_depths(gene_depths)
        self.set_y(y)

    def set_gene_depths(self, gene_depths):
        d = np.ascontiguousarray(gene_depths, dtype=np.float64)
        if d.ndim != 1:
            raise ValueError("gene_depths must be a vector with one entry per gene")
        B.check(B.lib().bvg_fit_set_gene_depths(self._h, _dp(d), d.shape[0]))

    def set_y(self, y):
        if y.ndim != 2 or y.shape[0] != self.M:
            raise ValueError(f"y must be M x G with M = {self.M}")
        self.G = y.shape[1]
        if np.issubdtype(y.dtype, np.integer):
            y = np.asfortranarray(y, dtype=np.int32)
            B.check(B.lib().bvg_fit_set_y_i32(self._h, y.ctypes.data_as(C.POINTER(C.c_int32)), self.M, self.G))
        else:
            y = np.asfortranarray(y, dtype=np.float64)
            B.check(B.lib().bvg_fit_set_y_f64(self._h, _dp(y), self.M, self.G))
        d = C.c_int64()
        B.check(B.lib().bvg_fit_dim(self._h, C.byref(d)))
        self.D = d.value

    def close(self):
        if getattr(self, "_h", None):
            B.lib().bvg_fit_destroy(self._h)
            self._h = None

    __del__ = close

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, len(unique_id))
        B.check(B.lib().bvg_fit_comm_init(self._h, C.cast(buf, C.c_void_p), len(unique_id)))

    def peer_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        n = C.c_size_t()
        B.check(B.lib().bvg_fit_peer_handle(self._h, C.cast(buf, C.c_void_p), 64, C.byref(n)))
        return buf.raw[: n.value]

    def peer_init(self, handles: bytes):
        buf = C.create_string_buffer(handles, len(handles))
        B.check(B.lib().bvg_fit_peer_init(self._h, C.cast(buf, C.c_void_p), len(handles)))

    def log_prob(self, theta, jacobian=True, propto=True, grad=True):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        if th.shape != (self.D,):
            raise ValueError(f"theta must have length {self.D}")
        lp = C.c_double()
        g = np.empty(self.D) if grad else None
        B.check(B.lib().bvg_log_prob(self._h, _dp(th), int(jacobian), int(propto), C.byref(lp), _dp(g) if grad else None))
        return (lp.value, g) if grad else lp.value

    def run(self):
        B.check(B.lib().bvg_fit_run(self._h))
        return self.summary()

    def mle(self, theta0=None):
        out = np.empty(self.D)
        t0 = None if theta0 is None else _dp(np.ascontiguousarray(theta0, dtype=np.float64))
        B.check(B.lib().bvg_fit_mle(self._h, t0, _dp(out)))
        return out

    def advi(self, theta_init=None):
        t0 = None if theta_init is None else _dp(np.ascontiguousarray(theta_init, dtype=np.float64))
        B.check(B.lib().bvg_fit_advi(self._h, t0))

    def summarize(self):
        B.check(B.lib().bvg_fit_summarize(self._h))
        return self.summary()

    def set_variational(self, mu, omega=None):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        om = None if omega is None else np.ascontiguousarray(omega, dtype=np.float64)
        B.check(B.lib().bvg_fit_set_variational(self._h, _dp(mu), None if om is None else _dp(om)))

    def get_variational(self):
        mu, om = np.empty(self.D), np.empty(self.D)
        B.check(B.lib().bvg_fit_get_variational(self._h, _dp(mu), _dp(om)))
        return mu, om

    def pathfinder_single(self, path=0):
        """algorithm = "pathfinder": one path -> (amplitude draws [pf_single_draws][G], lp - log q per draw, info)"""
        nd, it = self.cfg.pf_single_draws, self.cfg.pf_max_lbfgs_iters
        amp, lr, info, elbos = np.empty((nd, self.G)), np.empty(nd), np.empty(5), np.empty(it + 1)
        B.check(B.lib().bvg_fit_pathfinder_single(self._h, int(path), _dp(amp), _dp(lr), _dp(info), _dp(elbos)))
        return amp, lr, dict(lbfgs_iters=int(info[0]), lbfgs_term=int(info[1]), best_iter=int(info[2]), n_pairs=int(info[3]),
                             best_elbo=info[4], elbos=elbos)

    def get_cholesky(self):
        """algorithm = "fullrank": the packed row-major lower-triangular factor L of q = N(mu, L L^T)"""
        out = np.empty(self.D * (self.D + 1) // 2)
        B.check(B.lib().bvg_fit_get_cholesky(self._h, _dp(out)))
        return out

    def set_cholesky(self, L):
        L = np.ascontiguousarray(L, dtype=np.float64)
        if L.shape != (self.D * (self.D + 1) // 2,):
            raise ValueError("L must be the packed lower triangle, D (D + 1) / 2 entries")
        B.check(B.lib().bvg_fit_set_cholesky(self._h, _dp(L)))

    def advi_steps(self, n, eta, reset=False):
        ms = C.c_float()
        B.check(B.lib().bvg_fit_advi_steps(self._h, n, float(eta), int(reset), C.byref(ms)))
        return ms.value

    def elbo(self, n_samples=150):
        e, ms = C.c_double(), C.c_float()
        B.check(B.lib().bvg_fit_elbo(self._h, n_samples, C.byref(e), C.byref(ms)))
        return e.value, ms.value

    def time_lik_kernel(self, n=10, flush_l2=True, backward=True):
        ms = C.c_float()
        B.check(B.lib().bvg_fit_time_lik_kernel(self._h, n, int(flush_l2), int(backward), C.byref(ms)))
        return ms.value

    def stage_times(self, n=20):
        """diagnostics: average device ms of (k_prep, likelihood, k_gene + finish) run back to back"""
        out = (C.c_float * 4)()
        B.check(B.lib().bvg_debug_stage_times(self._h, n, out))
        return list(out)[:3]

    def tc_trace(self, nchunk=31):
        """diagnostics: clock64 stamps of CTA (0,0) of one backward launch of the tcgen05 kernel, [chunk][8]:
        0 fwd issuer loop top, 1 after MMA1 issue, 2 after MMA2 issue | epilogue thread 96: 3 Y landed, 4 D1 ready,
        5 after tcgen05.ld, 6 after the residual math, 7 residual published.  Row 30: kernel start / after setup /
        before final sync / end."""
        import numpy as np
        out = (C.c_longlong * (8 * nchunk))()
        B.check(B.lib().bvg_debug_tc_trace(self._h, out, nchunk))
        return np.array(list(out), dtype=np.int64).reshape(nchunk, 8)

    def persist_trace(self, n_iter=20):
        """diagnostics: clock64 stamps of CTA 0 of one k_advi_persist stretch, [iteration][8]: 0 iteration start, 1 passes done
        (partials published), 2 gene phases done, 3 grid barrier passed, 4 sums reduced (+ exchanged), 5 shared update done,
        6 state loaded (before the tile wait), 7 tile counter reached, 8-15 finer stamps of the gene phase / block row.  Second
        array: per-chunk stamps of the pass of iteration 8 (layout of tc_trace); row 31: coefficient rows stored, chunk loop
        done, D2 read, partials written."""
        out = (C.c_longlong * (32 * 16 + 32 * 8 + 32 * 4))()
        B.check(B.lib().bvg_debug_persist_trace(self._h, int(n_iter), out, 64))
        v = np.array(list(out), dtype=np.int64)
        self.last_latency_probe = v[768:].reshape(32, 4)
        return v[:512].reshape(32, 16), v[512:768].reshape(32, 8)

    def time_gene(self, n=100, mode=1, fuse=3, ahead=0):
        """diagnostics: average device ms of n back-to-back k_gene launches"""
        ms = C.c_float()
        B.check(B.lib().bvg_debug_time_gene(self._h, n, mode, fuse, ahead, C.byref(ms)))
        return ms.value

    def gene_trace(self, ahead=0):
        """diagnostics: SM cycle-counter stamps inside one k_gene launch.  Stamps 0-3 and 8-15 come from CTA 0
        (relative to stamp 0); stamps 4-7 from the last CTA to retire, on another SM (relative to stamp 4)."""
        out = (C.c_longlong * 16)()
        B.check(B.lib().bvg_debug_gene_trace(self._h, out, int(ahead)))
        v = list(out)
        return [(x - (v[4] if 4 <= i <= 7 else v[0])) if x else None for i, x in enumerate(v)]

    def flush_l2(self):
        B.check(B.lib().bvg_fit_flush_l2(self._h))

    def summary(self):
        out = np.empty((self.G, 8), order="F")
        B.check(B.lib().bvg_fit_get_summary(self._h, _dp(out)))
        return out

    def draws(self):
        out = np.empty((self.cfg.n_draws, self.G), order="F")
        B.check(B.lib().bvg_fit_get_draws(self._h, _dp(out)))
        return out

    def posterior_columns(self):
        """CmdStan's variational CSV header for this model: lp__, log_p__, log_g__, the parameters in declaration
        order (inst/approxGP*.stan `parameters`), then the transformed parameters (approxGP.stan:25-28)."""
        cols = posterior_column_names(self.cfg.likelihood == B.NB, bool(self.cfg.depth_adjust), self.G, self.k)
        assert len(cols) == 3 + self.D + (1 if self.cfg.depth_adjust else 2) * self.G
        return cols

    def posterior(self, n=None):
        """n x ncol draws of the whole model (what cmdstanr's `fit_vi$draws()` holds; save.model = TRUE, :483-489)."""
        n = int(self.cfg.n_draws if n is None else n)
        nc = C.c_int64()
        B.check(B.lib().bvg_fit_posterior_ncol(self._h, C.byref(nc)))
        out = np.empty((n, nc.value), order="F")
        B.check(B.lib().bvg_fit_get_posterior(self._h, n, _dp(out)))
        return out

    def trace(self):
        out = np.zeros((64, 4))
        n = C.c_int()
        B.check(B.lib().bvg_fit_get_trace(self._h, _dp(out), 64, C.byref(n)))
        return out[: n.value].copy()

    def stats(self):
        st = B.Stats()
        B.check(B.lib().bvg_fit_get_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in B.Stats._fields_ if f != "pad"}


def comm_unique_id():
    buf = C.create_string_buffer(128)
    n = C.c_size_t()
    B.check(B.lib().bvg_comm_unique_id(C.cast(buf, C.c_void_p), 128, C.byref(n)))
    return buf.raw[: n.value]

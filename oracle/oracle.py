"""ctypes binding of the CPU oracle (oracle/libbvg_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (bayesvg_b200/) never does.  PARITY UNPINNED -- see bvg_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbvg_oracle.so")

GAUSSIAN, NB = 0, 1
KERNELS = {"exp_quad": 0, "matern": 1, "periodic": 2}


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("bvg_oracle.c", "bvg_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


class _AdviCfg(C.Structure):
    _fields_ = [("iter", C.c_int), ("elbo_samples", C.c_int), ("eval_elbo", C.c_int),
                ("adapt_iter", C.c_int), ("tol_rel_obj", C.c_double), ("eta", C.c_double),
                ("seed", C.c_uint64)]


class _AdviRes(C.Structure):
    _fields_ = [("eta", C.c_double), ("iters", C.c_int), ("grad_evals", C.c_int),
                ("elbo_evals", C.c_int), ("converged", C.c_int), ("elbo", C.c_double),
                ("n_trace", C.c_int)]


class _PfCfg(C.Structure):
    _fields_ = [("max_lbfgs_iters", C.c_int), ("history", C.c_int), ("num_elbo_draws", C.c_int), ("num_draws", C.c_int),
                ("num_paths", C.c_int), ("num_psis_draws", C.c_int), ("init_radius", C.c_double), ("seed", C.c_uint64)]


class _PfInfo(C.Structure):
    _fields_ = [("lbfgs_iters", C.c_int), ("lbfgs_term", C.c_int), ("best_iter", C.c_int), ("n_pairs", C.c_int),
                ("best_elbo", C.c_double)]


class _LbfgsCfg(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("history", C.c_int), ("init_alpha", C.c_double),
                ("tol_obj", C.c_double), ("tol_rel_obj", C.c_double), ("tol_grad", C.c_double),
                ("tol_rel_grad", C.c_double), ("tol_param", C.c_double)]


class _LbfgsRes(C.Structure):
    _fields_ = [("iters", C.c_int), ("fevals", C.c_int), ("term", C.c_int), ("lp", C.c_double)]


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.bvgo_create.restype = C.c_void_p
        L.bvgo_create.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int, _dp, _dp, C.c_int]
        L.bvgo_destroy.argtypes = [C.c_void_p]
        L.bvgo_set_depths.argtypes = [C.c_void_p, _dp]
        L.bvgo_set_depths.restype = None
        L.bvgo_dim.restype = C.c_int64
        L.bvgo_dim.argtypes = [C.c_int, C.c_int64, C.c_int]
        L.bvgo_log_prob.restype = C.c_double
        L.bvgo_log_prob.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp]
        L.bvgo_normal.restype = C.c_double
        L.bvgo_normal.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
        L.bvgo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.bvgo_elbo.restype = C.c_double
        L.bvgo_elbo.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_uint64, C.c_uint32]
        L.bvgo_advi_step.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_int, C.c_uint64, C.c_uint32]
        L.bvgo_advi.argtypes = [C.c_void_p, _dp, C.POINTER(_AdviCfg), _dp, _dp, C.POINTER(_AdviRes), _dp, C.c_int]
        L.bvgo_lbfgs.argtypes = [C.c_void_p, _dp, C.POINTER(_LbfgsCfg), _dp, C.POINTER(_LbfgsRes)]
        L.bvgo_summary.argtypes = [C.c_int, C.c_int64, C.c_int, _dp, _dp, C.c_int, C.c_uint64, C.c_int64, C.c_int64, _dp]
        L.bvgo_basis.argtypes = [_dp, C.c_int64, _dp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, _dp, _dp]
        L.bvgo_fr_elbo.restype = C.c_double
        L.bvgo_fr_elbo.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_uint64, C.c_uint32]
        L.bvgo_fr_step.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_int, C.c_uint64, C.c_uint32]
        L.bvgo_advi_fullrank.argtypes = L.bvgo_advi.argtypes
        L.bvgo_summary_fullrank.argtypes = [C.c_int, C.c_int64, C.c_int, _dp, _dp, C.c_int, C.c_uint64, _dp]
        L.bvgo_pf_single.argtypes = [C.c_void_p, C.POINTER(_PfCfg), C.c_int, _dp, _dp, C.POINTER(_PfInfo), _dp]
        L.bvgo_pf_single_phi.argtypes = [C.c_void_p, C.POINTER(_PfCfg), C.c_int, _dp, _dp, C.POINTER(_PfInfo), _dp]
        L.bvgo_psis_weights.restype = C.c_double
        L.bvgo_psis_weights.argtypes = [_dp, C.c_int, _dp]
        L.bvgo_pathfinder.argtypes = [C.c_void_p, C.POINTER(_PfCfg), _dp, C.POINTER(_PfInfo), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.bvgo_uniform.restype = C.c_double
        L.bvgo_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
        L.bvgo_kmeans.argtypes = [_dp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64, _dp, _dp, C.POINTER(C.c_int32)]
        L.bvgo_digamma.restype = C.c_double
        L.bvgo_digamma.argtypes = [C.c_double]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


class Oracle:
    """One dataset: Y (M x G), Phi (M x k).  model: 'gaussian' | 'nb'; gene_depths (length G) selects the
    depth-adjusted programs inst/approxGP2.stan / approxGP3.stan."""

    def __init__(self, phi, y, model="gaussian", nthreads=1, gene_depths=None):
        self.model = {"gaussian": GAUSSIAN, "nb": NB}[model] + (0 if gene_depths is None else 2)
        self.phi, self.y = _f(phi), _f(y)
        self.M, self.k = self.phi.shape
        assert self.y.shape[0] == self.M
        self.G = self.y.shape[1]
        self.D = int(lib().bvgo_dim(self.model, self.G, self.k))
        self._h = lib().bvgo_create(self.model, self.M, self.G, self.k, _p(self.phi), _p(self.y), nthreads)
        if gene_depths is not None:
            self.depths = np.ascontiguousarray(gene_depths, dtype=np.float64)
            assert self.depths.shape == (self.G,)
            lib().bvgo_set_depths(self._h, _p(self.depths))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().bvgo_destroy(self._h)
            self._h = None

    def log_prob(self, theta, jacobian=True, propto=True, grad=True):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        assert th.shape == (self.D,)
        g = np.empty(self.D) if grad else None
        lp = lib().bvgo_log_prob(self._h, _p(th), int(jacobian), int(propto), _p(g) if grad else None)
        return (lp, g) if grad else lp

    def elbo(self, mu, omega, S=150, seed=312, t0=0):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        om = np.ascontiguousarray(omega, dtype=np.float64)
        return lib().bvgo_elbo(self._h, _p(mu), _p(om), S, seed, t0)

    def advi_step(self, mu, omega, hist, eta, it, seed, t):
        """in-place on mu, omega (D) and hist (2D)"""
        return lib().bvgo_advi_step(self._h, _p(mu), _p(omega), _p(hist), eta, it, seed, t)

    def advi(self, theta_init, iter=3000, elbo_samples=150, eval_elbo=100, adapt_iter=50,
             tol_rel_obj=0.01, eta=0.0, seed=312):
        cfg = _AdviCfg(iter, elbo_samples, eval_elbo, adapt_iter, tol_rel_obj, eta, seed)
        res = _AdviRes()
        th = np.ascontiguousarray(theta_init, dtype=np.float64)
        mu, om = np.empty(self.D), np.empty(self.D)
        cap = iter // eval_elbo + 1
        tr = np.zeros((cap, 4))
        rc = lib().bvgo_advi(self._h, _p(th), C.byref(cfg), _p(mu), _p(om), C.byref(res), _p(tr), cap)
        info = {f: getattr(res, f) for f, _ in _AdviRes._fields_}
        info["rc"] = rc
        info["trace"] = tr[: res.n_trace].copy()
        return mu, om, info

    # ---- algorithm = "fullrank": L is the PACKED row-major lower-triangular Cholesky factor, D (D + 1) / 2 entries ----
    def fr_elbo(self, mu, L, S=150, seed=312, t0=0):
        mu, L = np.ascontiguousarray(mu, dtype=np.float64), np.ascontiguousarray(L, dtype=np.float64)
        assert L.shape == (self.D * (self.D + 1) // 2,)
        return lib().bvgo_fr_elbo(self._h, _p(mu), _p(L), S, seed, t0)

    def fr_step(self, mu, L, hist, eta, it, seed, t):
        """in-place on mu (D), L (packed) and hist (D + D (D + 1) / 2)"""
        assert hist.shape == (self.D + self.D * (self.D + 1) // 2,)
        return lib().bvgo_fr_step(self._h, _p(mu), _p(L), _p(hist), eta, it, seed, t)

    def advi_fullrank(self, theta_init, iter=3000, elbo_samples=150, eval_elbo=100, adapt_iter=50,
                      tol_rel_obj=0.01, eta=0.0, seed=312):
        cfg = _AdviCfg(iter, elbo_samples, eval_elbo, adapt_iter, tol_rel_obj, eta, seed)
        res = _AdviRes()
        th = np.ascontiguousarray(theta_init, dtype=np.float64)
        mu, L = np.empty(self.D), np.empty(self.D * (self.D + 1) // 2)
        cap = iter // eval_elbo + 1
        tr = np.zeros((cap, 4))
        rc = lib().bvgo_advi_fullrank(self._h, _p(th), C.byref(cfg), _p(mu), _p(L), C.byref(res), _p(tr), cap)
        info = {f: getattr(res, f) for f, _ in _AdviRes._fields_}
        info["rc"] = rc
        info["trace"] = tr[: res.n_trace].copy()
        return mu, L, info

    def summary_fullrank(self, mu, L, n_draws=1000, seed=312):
        mu, L = np.ascontiguousarray(mu, dtype=np.float64), np.ascontiguousarray(L, dtype=np.float64)
        out = np.empty((self.G, 8), order="F")
        lib().bvgo_summary_fullrank(self.model, self.G, self.k, _p(mu), _p(L), n_draws, seed, _p(out))
        return out

    # ---- algorithm = "pathfinder" ----
    @staticmethod
    def _pf_cfg(max_lbfgs_iters=200, history=25, num_elbo_draws=50, num_draws=1000, num_paths=2, num_psis_draws=1000,
                init_radius=0.01, seed=312):
        return _PfCfg(max_lbfgs_iters, history, num_elbo_draws, num_draws, num_paths, num_psis_draws, init_radius, seed)

    def pf_single(self, path=0, **kw):
        cfg = self._pf_cfg(**kw)
        amp = np.empty((cfg.num_draws, self.G))
        lr = np.empty(cfg.num_draws)
        elbos = np.empty(cfg.max_lbfgs_iters + 1)
        info = _PfInfo()
        rc = lib().bvgo_pf_single(self._h, C.byref(cfg), path, _p(amp), _p(lr), C.byref(info), _p(elbos))
        d = {f: getattr(info, f) for f, _ in _PfInfo._fields_}
        d.update(rc=rc, elbos=elbos)
        return amp, lr, d

    def pf_single_phi(self, path=0, **kw):
        """one Pathfinder path with its draws in the unconstrained space: (phi [num_draws][D], lp - log q, info)"""
        cfg = self._pf_cfg(**kw)
        phi = np.empty((cfg.num_draws, self.D))
        lr = np.empty(cfg.num_draws)
        elbos = np.empty(cfg.max_lbfgs_iters + 1)
        info = _PfInfo()
        rc = lib().bvgo_pf_single_phi(self._h, C.byref(cfg), path, _p(phi), _p(lr), C.byref(info), _p(elbos))
        d = {f: getattr(info, f) for f, _ in _PfInfo._fields_}
        d.update(rc=rc, elbos=elbos)
        return phi, lr, d

    def psis_weights(self, lr):
        lr = np.ascontiguousarray(lr, dtype=np.float64)
        w = np.empty_like(lr)
        lib().bvgo_psis_weights.restype = C.c_double
        kh = lib().bvgo_psis_weights(_p(lr), lr.shape[0], _p(w))
        return w, kh

    def pathfinder(self, **kw):
        cfg = self._pf_cfg(**kw)
        out = np.empty((self.G, 8), order="F")
        infos = (_PfInfo * cfg.num_paths)()
        idx = (C.c_int * cfg.num_psis_draws)()
        kh = C.c_double()
        rc = lib().bvgo_pathfinder(self._h, C.byref(cfg), _p(out), infos, idx, C.byref(kh))
        return out, dict(rc=rc, khat=kh.value, idx=np.array(list(idx)),
                         paths=[{f: getattr(i, f) for f, _ in _PfInfo._fields_} for i in infos])

    def lbfgs(self, theta0=None, max_iter=1000, history=25):
        th0 = np.zeros(self.D) if theta0 is None else np.ascontiguousarray(theta0, dtype=np.float64)
        cfg = _LbfgsCfg(max_iter, history, 1e-3, 1e-12, 1e4, 1e-8, 1e7, 1e-8)
        res = _LbfgsRes()
        out = np.empty(self.D)
        rc = lib().bvgo_lbfgs(self._h, _p(th0), C.byref(cfg), _p(out), C.byref(res))
        return out, {"rc": rc, "iters": res.iters, "fevals": res.fevals, "term": res.term, "lp": res.lp}

    def summary(self, mu, omega, n_draws=1000, seed=312, gene_offset=0, G_total=None):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        om = np.ascontiguousarray(omega, dtype=np.float64)
        out = np.empty((self.G, 8), order="F")
        lib().bvgo_summary(self.model, self.G, self.k, _p(mu), _p(om), n_draws, seed, gene_offset,
                           self.G if G_total is None else G_total, _p(out))
        return out

    def posterior(self, mu, omega, n, seed=312):
        """n draws of the whole model as CmdStan's variational CSV lays them out [U] (SURVEY A.4 "output"):
        lp__ = 0, log_p__ = log_prob<propto=false, jacobian=true>(zeta), log_g__ = -1/2 sum eps^2, the parameters
        constrained in declaration order, then beta0[G] (hierarchical intercept only; approxGP.stan:26) and
        amplitude_sq[G] (:27).  Draw d of parameter i uses Philox stream 3, counter d (as `summary`)."""
        depth = self.model >= 2
        lik = self.model & 1
        L = layout(lik, self.G, self.k, depth)
        G, k, D = self.G, self.k, self.D
        pos = np.zeros(D, bool)
        pos[L["sigma_amplitude"]] = True
        pos[L["amplitude"]:L["amplitude"] + G] = True
        pos[L["sigma_alpha"]:L["sigma_alpha"] + k] = True
        pos[L["phi_nb" if lik == NB else "sigma_y"]] = True
        if not depth:
            pos[L["sigma_beta0"]] = True
        out = np.empty((n, 3 + D + (1 if depth else 2) * G), order="F")
        for d in range(n):
            eps = np.array([normal(seed, 3, d, i) for i in range(D)])
            zeta = np.asarray(mu) + np.exp(omega) * eps
            th = np.where(pos, np.exp(zeta), zeta)
            out[d, 0] = 0.0
            out[d, 1] = self.log_prob(zeta, True, False, grad=False)
            out[d, 2] = -0.5 * float(eps @ eps)
            out[d, 3:3 + D] = th
            c = 3 + D
            if not depth:
                out[d, c:c + G] = th[L["mu_beta0"]] + th[L["sigma_beta0"]] * th[L["z_beta0"]:L["z_beta0"] + G]
                c += G
            out[d, c:c + G] = th[L["amplitude"]:L["amplitude"] + G] ** 2
        return out


def pack_tril(Lfull):
    """dense lower-triangular D x D -> packed row-major (row i at i (i + 1) / 2)"""
    D = Lfull.shape[0]
    return np.ascontiguousarray(Lfull[np.tril_indices(D)])


def unpack_tril(Lp, D):
    out = np.zeros((D, D))
    out[np.tril_indices(D)] = Lp
    return out


class GmmOracle(Oracle):
    """inst/clusterSVGs.stan on the PCA embedding X (N x D): log_prob / advi / advi_fullrank are inherited"""

    def __init__(self, X, K):
        self.X = np.ascontiguousarray(X, dtype=np.float64)   # row-major [N][D]
        self.N, self.Dx = self.X.shape
        self.K = int(K)
        self.G, self.k, self.model = self.K, 0, 4
        lib().bvgo_gmm_create.restype = C.c_void_p
        lib().bvgo_gmm_create.argtypes = [_dp, C.c_int64, C.c_int, C.c_int]
        lib().bvgo_gmm_post.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_uint64, _dp, _dp, C.POINTER(C.c_double)]
        self._h = lib().bvgo_gmm_create(_p(self.X), self.N, self.Dx, self.K)
        self.D = self.K - 1 + 2 * self.K * self.Dx

    def post(self, mu, var, fullrank, n_draws=1000, seed=312):
        mu, var = np.ascontiguousarray(mu, dtype=np.float64), np.ascontiguousarray(var, dtype=np.float64)
        draws = np.empty((n_draws, self.K + 2 * self.K * self.Dx))
        soft = np.empty((self.N, self.K))
        ll = C.c_double()
        lib().bvgo_gmm_post(self._h, _p(mu), _p(var), int(fullrank), n_draws, seed, _p(draws), _p(soft), C.byref(ll))
        return draws, soft, ll.value


def variogram(coords, expr, kappa=1.5):
    """lscale.estimator = "variogram" (findSpatiallyVariableFeaturesBayes.R:208-260), numpy / scipy restatement for small cases:
    gstat::variogram(z ~ 1) defaults [U] (cutoff = bounding-box diagonal / 3, 15 classes of width cutoff / 15, gamma = sum (z_i -
    z_j)^2 / (2 N_h), dist = mean pair distance) and gstat::fit.variogram of nugget + Matern(kappa) by weighted least squares with
    weights N_h / h^2 [U], started at vgm(0.8 var, "Mat", median(dist), 0.2 var) (:241-245).  Returns (median range, ranges, table)."""
    from scipy.optimize import least_squares
    X, Z = np.asarray(coords, dtype=np.float64), np.asarray(expr, dtype=np.float64)
    M, G = X.shape[0], Z.shape[0]
    cutoff = np.sqrt(((X.max(0) - X.min(0)) ** 2).sum()) / 3.0
    width = cutoff / 15
    iu, ju = np.triu_indices(M, 1)
    dx, dy = X[iu, 0] - X[ju, 0], X[iu, 1] - X[ju, 1]
    h = np.sqrt(dx * dx + dy * dy)
    keep = (h > 0) & (h <= cutoff)
    iu, ju, h = iu[keep], ju[keep], h[keep]
    b = np.minimum(np.floor(h / width).astype(int), 14)
    npairs = np.bincount(b, minlength=15).astype(np.float64)
    dist = np.bincount(b, weights=h, minlength=15) / np.maximum(npairs, 1)
    use = npairs > 0
    rho = {0.5: lambda x: np.exp(-x), 1.5: lambda x: (1 + x) * np.exp(-x), 2.5: lambda x: (1 + x + x * x / 3) * np.exp(-x)}[kappa]
    ranges, gam_all = np.full(G, np.nan), np.zeros((G, 15))
    for g in range(G):
        d2 = (Z[g, iu] - Z[g, ju]) ** 2
        gam = np.bincount(b, weights=d2, minlength=15) / (2 * np.maximum(npairs, 1))
        gam_all[g] = gam
        var = Z[g].var(ddof=1)
        if not var > 0 or use.sum() < 3:
            continue
        hh, gg, ww = dist[use], gam[use], np.sqrt(npairs[use] / dist[use] ** 2)
        p0 = np.array([0.2 * var, 0.8 * var, np.median(hh)])
        res = least_squares(lambda p: ww * (gg - (p[0] + p[1] * (1 - rho(hh / p[2])))) if p[2] > 0 else np.full(len(hh), 1e150), p0,
                            method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-14, max_nfev=2000)
        if np.isfinite(res.x[2]) and res.x[2] > 0:
            ranges[g] = res.x[2]
    return float(np.nanmedian(ranges)), ranges, dict(np=npairs, dist=dist, gamma=gam_all, cutoff=cutoff)


def pf_approx(alpha, S, Y, x, g, u):
    """test hook: Pathfinder's Gaussian approximation from curvature pairs S, Y ([n][D], oldest first)"""
    alpha, S, Y, x, g, u = (np.ascontiguousarray(a, dtype=np.float64) for a in (alpha, S, Y, x, g, u))
    D, n, nu = len(alpha), S.shape[0], u.shape[0]
    phi, lq, xc = np.empty((nu, D)), np.empty(nu), np.empty(D)
    lib().bvgo_pf_approx_test.argtypes = [C.c_int64, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp]
    rc = lib().bvgo_pf_approx_test(D, n, _p(alpha), _p(S), _p(Y), _p(x), _p(g), nu, _p(u), _p(phi), _p(lq), _p(xc))
    assert rc == 0
    return phi, lq, xc


def psis_weights(log_ratios):
    lr = np.ascontiguousarray(log_ratios, dtype=np.float64)
    w = np.empty(len(lr))
    k = lib().bvgo_psis_weights(_p(lr), len(lr), _p(w))
    return w, k


def uniform(seed, stream, t, i):
    return lib().bvgo_uniform(seed, stream, t, i)


def normal(seed, stream, t, i):
    return lib().bvgo_normal(seed, stream, t, i)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().bvgo_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def basis(coords, centers, lscale, kernel="matern", nu=2.5, period=100.0, return_raw=False):
    X, Cc = _f(coords), _f(centers)
    M, k = X.shape[0], Cc.shape[0]
    Q = np.empty((M, k), order="F")
    raw = np.empty((M, k), order="F") if return_raw else None
    lib().bvgo_basis(_p(X), M, _p(Cc), k, lscale, KERNELS[kernel], nu, period, _p(Q), _p(raw) if return_raw else None)
    return (Q, raw) if return_raw else Q


def kmeans(coords, k, iter_max=100, nstart=10, seed=312):
    X = _f(coords)
    out = np.empty((k, 2), order="F")
    ss = C.c_double()
    info = (C.c_int32 * 3)()
    rc = lib().bvgo_kmeans(_p(X), X.shape[0], k, iter_max, nstart, seed, _p(out), C.byref(ss), info)
    assert rc == 0
    return out, dict(tot_withinss=ss.value, start=info[0], iters=info[1], converged=bool(info[2]))


def digamma(x):
    return lib().bvgo_digamma(x)


def layout(model, G, k, depth=False):
    """slot offsets of the unconstrained vector"""
    if depth:  # inst/approxGP2.stan:13-23, inst/approxGP3.stan:13-23
        o = dict(beta0=0, beta1=1, amplitude=3, mu_amplitude=3 + G, sigma_amplitude=4 + G, mu_alpha=5 + G,
                 sigma_alpha=5 + G + k, z_alpha_t=5 + G + 2 * k, D=5 + G + 2 * k + k * G)
        o["sigma_y" if model in ("gaussian", GAUSSIAN) else "phi_nb"] = 2
        return o
    if model in ("gaussian", GAUSSIAN):
        o = dict(mu_beta0=0, sigma_beta0=1, z_beta0=2, mu_amplitude=2 + G, sigma_amplitude=3 + G,
                 amplitude=4 + G, mu_alpha=4 + 2 * G, sigma_alpha=4 + 2 * G + k, z_alpha_t=4 + 2 * G + 2 * k,
                 sigma_y=4 + 2 * G + 2 * k + k * G)
    else:
        o = dict(mu_beta0=0, sigma_beta0=1, z_beta0=2, phi_nb=2 + G, amplitude=3 + G, mu_amplitude=3 + 2 * G,
                 sigma_amplitude=4 + 2 * G, mu_alpha=5 + 2 * G, sigma_alpha=5 + 2 * G + k,
                 z_alpha_t=5 + 2 * G + 2 * k)
    o["D"] = 5 + 2 * G + 2 * k + k * G
    return o

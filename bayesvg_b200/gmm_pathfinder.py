"""algorithm = "pathfinder" of clusterSVGsBayes() (R/clusterSVGsBayes.R:162-170):

    mod$pathfinder(data_list, seed, init = 0.01, draws = n.draws, num_elbo_draws = elbo.samples,
                   max_lbfgs_iters = 500L, history_size = 100L)                       [num_paths = 4: CmdStan's default]

Pathfinder [U: stan::services::pathfinder::pathfinder_lbfgs_single / _multi; Zhang, Carpenter, Gelman, Vehtari 2022] as the oracle
restates it for the main model (oracle/bvg_oracle.c: pf_build / pf_draw / psis), here for inst/clusterSVGs.stan (P = K - 1 +
2 K D = 304 parameters at the defaults).  The split of work is CmdStan's: the host walks the L-BFGS path and does the (2 J)-
dimensional algebra of every iterate's Gaussian approximation (numpy / LAPACK; J = 100), and EVERY log density and gradient --
the L-BFGS evaluations and the num_elbo_draws x iterates x paths ELBO draws -- is evaluated on the device with the embedding
resident there (bvg_gmm_open / bvg_gmm_eval: one launch per iterate, one CTA per draw).  There is no CPU evaluator.

RNG: the library's counter-based generator (Philox4x32-10 + Box-Muller), streams 5-8 as in the oracle:
init (5), ELBO draws (6), path draws (7), PSIS resampling (8).
"""
import ctypes as C

import numpy as np

from . import _lib as B

STREAM_PF_INIT, STREAM_PF_ELBO, STREAM_PF_DRAW, STREAM_PF_PSIS = 5, 6, 7, 8
_M0, _M1, _W0, _W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_LOG_TWO_PI = 1.8378770664093454835606594728112


def _philox(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint32 arrays (bvg_internal.cuh: philox4x32_10)"""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = p1.astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = p0.astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def _counters(seed, stream, t, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    n = idx.shape
    return ((idx & np.uint64(0xFFFFFFFF)).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32), np.full(n, t, dtype=np.uint32),
            np.full(n, stream, dtype=np.uint32), np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))


def uniform(seed, stream, t, idx):
    """bvgo_uniform / the u1 of bvg_normal: 53 bits of (o[0], o[2])"""
    o0, _, o2, _ = _philox(*_counters(seed, stream, t, idx))
    return (((o0.astype(np.uint64) << np.uint64(21)) | (o2.astype(np.uint64) >> np.uint64(11))).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def normal(seed, stream, t, idx):
    """bvg_normal (bvg_internal.cuh): Box-Muller on (u1 from o[0], o[2]; u2 from o[1])"""
    o0, o1, o2, _ = _philox(*_counters(seed, stream, t, idx))
    u1 = (((o0.astype(np.uint64) << np.uint64(21)) | (o2.astype(np.uint64) >> np.uint64(11))).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
    u2 = (o1.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586476925286766559 * u2)


class DeviceEvaluator:
    """log_prob<propto = true, jacobian = true> of inst/clusterSVGs.stan on the device (bvg_gmm_open / bvg_gmm_eval)"""

    def __init__(self, X, K, device=0):
        self.X = np.asfortranarray(X, dtype=np.float64)
        self.N, self.D = self.X.shape
        self.K = int(K)
        self.P = self.K - 1 + 2 * self.K * self.D
        self._h = C.c_void_p()
        B.check(B.lib().bvg_gmm_open(device, self.X.ctypes.data_as(C.POINTER(C.c_double)), self.N, self.D, self.K, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            B.lib().bvg_gmm_close(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def lp_grad(self, theta):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        lp, g = np.empty(1), np.empty(self.P)
        B.check(B.lib().bvg_gmm_eval(self._h, _dp(th), 1, 1, 1, _dp(lp), _dp(g)))
        return float(lp[0]), g

    def lp_batch(self, thetas):
        th = np.ascontiguousarray(thetas, dtype=np.float64)
        lp = np.empty(th.shape[0])
        B.check(B.lib().bvg_gmm_eval(self._h, _dp(th), th.shape[0], 1, 1, _dp(lp), None))
        return lp

    def posterior(self, thetas, want_draws=True):
        th = np.ascontiguousarray(thetas, dtype=np.float64)
        nd = th.shape[0]
        draws = np.empty((nd, self.K + 2 * self.K * self.D)) if want_draws else None
        soft, ll = np.empty((self.N, self.K)), C.c_double()
        B.check(B.lib().bvg_gmm_posterior(self._h, _dp(th), nd, _dp(draws) if want_draws else None, _dp(soft), C.byref(ll)))
        return draws, soft, ll.value


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ---- L-BFGS: CmdStan's BFGSMinimizer<LBFGSUpdate> with its Wolfe line search (the host form of csrc/lbfgs.cu) ------------------
def _cubic_interp(df0, x1, f1, df1, lo, hi):
    with np.errstate(all="ignore"):
        c3 = (-12 * f1 + 6 * x1 * (df0 + df1)) / (x1 * x1 * x1)
        c2 = -(4 * df0 + 2 * df1) / x1 + 6 * f1 / (x1 * x1)
        c1 = df0
        ts = np.sqrt(c2 * c2 - 2.0 * c1 * c3)
        r1, r2 = -(c2 + ts) / c3, -(c2 - ts) / c3

    def val(x):
        return x * (x * (x * c3 / 3.0 + c2) / 2.0 + c1)

    bx, bf = lo, val(lo)
    for cand in (hi, r1, r2):
        if cand == hi or (lo < cand < hi):
            v = val(cand)
            if v < bf:
                bf, bx = v, cand
    return bx


def _zoom(ev, f0, c1dfp, c2dfp, alo, aloF, aloD, ahi, ahiF, ahiD):
    itn = 1
    while True:
        if abs(alo - ahi) < 1e-16:
            return None
        if itn % 5 == 0:
            a = 0.5 * (alo + ahi)
        else:
            with np.errstate(all="ignore"):
                d1 = aloD + ahiD - 3 * (aloF - ahiF) / (alo - ahi)
                d2 = np.sqrt(d1 * d1 - aloD * ahiD)
                if ahi < alo:
                    d2 = -d2
                a = ahi - (ahi - alo) * (ahiD + d2 - d1) / (ahiD - aloD + 2 * d2)
            mn, mx, w = min(alo, ahi), max(alo, ahi), abs(alo - ahi)
            if not np.isfinite(a) or a < mn + 0.01 * w or a > mx - 0.01 * w:
                a = 0.5 * (alo + ahi)
        r = ev(a)
        while r is None:
            a = 0.5 * (a + min(alo, ahi))
            if abs(min(alo, ahi) - a) < 1e-16:
                return None
            r = ev(a)
        f1, dg1 = r[0], r[1]
        if f1 > f0 + a * c1dfp or f1 >= aloF:
            ahi, ahiF, ahiD = a, f1, dg1
        else:
            if abs(dg1) <= -c2dfp:
                return a, r
            if dg1 * (ahi - alo) >= 0:
                ahi, ahiF, ahiD = alo, aloF, aloD
            alo, aloF, aloD = a, f1, dg1
        itn += 1


def _wolfe(ev, alpha, f0, dfp):
    c1dfp, c2dfp = 1e-4 * dfp, 0.9 * dfp
    a0, prevF, prevD = 1e-12, f0, dfp
    nits = restarts = 0
    while True:
        if nits >= 20:
            return None
        r = ev(alpha)
        if r is None:
            if restarts >= 10:
                return None
            alpha = 0.5 * (a0 + alpha)
            restarts += 1
            continue
        restarts = 0
        f1, dg1 = r[0], r[1]
        if f1 > f0 + alpha * c1dfp or (f1 >= prevF and nits > 0):
            return _zoom(ev, f0, c1dfp, c2dfp, a0, prevF, prevD, alpha, f1, dg1)
        if abs(dg1) <= -c2dfp:
            return alpha, r
        if dg1 >= 0:
            return _zoom(ev, f0, c1dfp, c2dfp, alpha, f1, dg1, a0, prevF, prevD)
        a0, prevF, prevD = alpha, f1, dg1
        alpha *= 10.0
        nits += 1


def lbfgs_path(fg, x0, max_iter, history, on_iterate):
    """minimise f = -lp from x0; fg(x) -> (f, grad f) or None when not finite; on_iterate(it, x, g, f) at the initial point
    (it = 0) and at every accepted iterate.  Returns (iterations, termination code) with lbfgs.cu's codes."""
    eps, tol_obj, tol_rel_obj, tol_grad, tol_rel_grad, tol_param = np.finfo(float).eps, 1e-12, 1e4, 1e-8, 1e7, 1e-8
    r = fg(x0)
    if r is None:
        raise FloatingPointError("Pathfinder: log density is not finite at the initial point")
    fk, g = r
    x = x0.copy()
    on_iterate(0, x, g, fk)
    S, Y = [], []          # newest last
    p = -g
    dfp = float(g @ p)
    it, term = 0, 0
    fk1 = alphak1 = dfp_prev = 0.0
    alpha = 1e-3
    while not term and max_iter > 0:
        it += 1
        resetB = 1 if it == 1 else 0
        while True:
            if resetB:
                p = -g
                dfp = -float(g @ g)
            if it > 1 and resetB != 2:
                alpha = min(1.0, 1.01 * _cubic_interp(dfp_prev, alphak1, fk - fk1, dfp, 1e-12, 1.0))
            else:
                alpha = 1e-3

            def ev(a, x=x, p=p):
                rr = fg(x + a * p)
                if rr is None:
                    return None
                dg = float(rr[1] @ p)
                return (rr[0], dg, rr[1], a) if np.isfinite(dg) else None

            res = _wolfe(ev, alpha, fk, dfp)
            if res is None:
                if resetB:
                    term = -1
                    break
                resetB = 2
                continue
            break
        if term:
            break
        alpha, (f1, _, g1, a_used) = res
        x1 = x + a_used * p
        s, y = x1 - x, g1 - g
        x, g = x1, g1
        fk1, fk = fk, f1
        dfp_prev = dfp
        on_iterate(it, x, g, fk)
        if resetB:
            S, Y = [], []
        S.append(s)
        Y.append(y)
        if len(S) > history:
            S.pop(0)
            Y.pop(0)
        skyk, ykyk = float(s @ y), float(y @ y)
        snorm, gnorm = np.sqrt(float(s @ s)), np.sqrt(float(g @ g))
        if resetB:
            B0 = ykyk / skyk
            dfp_prev /= B0
            alphak1 = alpha * B0
        else:
            alphak1 = alpha
        gamma = skyk / ykyk
        if abs(fk1 - fk) < tol_obj:
            term = 1
        elif abs(fk1 - fk) < tol_rel_obj * eps * max(abs(fk1), abs(fk), 1.0):
            term = 2
        elif gnorm < tol_grad:
            term = 3
        elif snorm < tol_param:
            term = 4
        elif it >= max_iter:
            term = 5
        q = g.copy()
        al = []
        for sv, yv in zip(reversed(S), reversed(Y)):
            a_ = float(sv @ q) / float(sv @ yv)
            al.append(a_)
            q -= a_ * yv
        q *= gamma
        for (sv, yv), a_ in zip(zip(S, Y), reversed(al)):
            b_ = float(yv @ q) / float(sv @ yv)
            q += (a_ - b_) * sv
        p = -q
        dfp = float(g @ p)
        if not term and -dfp / max(abs(fk), 1.0) < tol_rel_grad * eps:
            term = 6
    return it, term


# ---- the Gaussian approximation of one iterate (oracle/bvg_oracle.c: pf_build, pf_draw) ----------------------------------------
class _Approx:
    __slots__ = ("n", "W", "xc", "alpha", "Lc", "La", "logdet")


def _build(alpha, S, Y, x, g):
    """S, Y: [n][P] oldest first; returns None when the history is linearly dependent"""
    from scipy.linalg import cholesky, solve_triangular
    n, P = len(S), x.shape[0]
    a = _Approx()
    a.n, a.alpha = n, alpha.copy()
    sq = np.sqrt(alpha)
    if n == 0:
        a.W, a.Lc, a.La = np.zeros((0, P)), np.zeros((0, 0)), np.zeros((0, 0))
        a.logdet = 0.5 * float(np.sum(np.log(alpha)))
        a.xc = x - alpha * g
        return a
    Sm, Ym = np.asarray(S), np.asarray(Y)
    SY = Sm @ Ym.T
    YaY = (Ym * alpha) @ Ym.T
    R = np.triu(SY)
    C_ = -solve_triangular(R, Sm, lower=False)          # C = -R^{-1} S
    W = np.vstack([Ym * sq, C_ / sq])
    Gm = W @ W.T
    try:
        Lc = cholesky(Gm, lower=True)
    except np.linalg.LinAlgError:
        return None
    if not np.all(np.isfinite(Lc)):
        return None
    Mk = np.zeros((2 * n, 2 * n))
    Mk[:n, n:] = np.eye(n)
    Mk[n:, :n] = np.eye(n)
    Mk[n:, n:] = YaY + np.diag(np.diag(SY))
    H = Lc.T @ Mk @ Lc + np.eye(2 * n)
    H = 0.5 * (H + H.T)
    try:
        La = cholesky(H, lower=True)
    except np.linalg.LinAlgError:
        return None
    if not np.all(np.isfinite(La)):
        return None
    a.W, a.Lc, a.La = W, Lc, La
    a.logdet = float(np.sum(np.log(np.diag(La)))) + 0.5 * float(np.sum(np.log(alpha)))
    t1 = (W[n:] * sq) @ g                                # C g
    t2 = (Ym * alpha) @ g
    v = t2 + (YaY + np.diag(np.diag(SY))) @ t1
    a.xc = x - (alpha * g + alpha * (Ym.T @ t1) + (W[n:] * sq).T @ v)
    return a


def _draw(a, U):
    """U: [B][P] standard normals -> (phi [B][P], log q [B])"""
    from scipy.linalg import solve_triangular
    P = U.shape[1]
    uu = np.sum(U * U, axis=1)
    if a.n == 0:
        phi = a.xc + np.sqrt(a.alpha) * U
    else:
        wu = a.W @ U.T                                   # [2n][B]
        u1 = solve_triangular(a.Lc, wu, lower=True)
        z = a.La @ u1 - u1
        cc = solve_triangular(a.Lc.T, z, lower=False)
        phi = a.xc + np.sqrt(a.alpha) * (U + (a.W.T @ cc).T)
    return phi, -a.logdet - 0.5 * (uu + P * _LOG_TWO_PI)


def single_path(ev, path, seed, num_draws, max_lbfgs_iters, history, num_elbo_draws, init_radius):
    """one Pathfinder path: (draws [num_draws][P] unconstrained, lp - log q [num_draws], info)"""
    P = ev.P
    idx = np.arange(P, dtype=np.uint64)
    st = dict(alpha=np.ones(P), S=[], Y=[], xprev=None, gprev=None, best=None, best_elbo=-np.inf, best_iter=-1, elbos=[])

    def fg(x):
        lp, g = ev.lp_grad(x)
        if not (np.isfinite(lp) and np.all(np.isfinite(g))):
            return None
        return -lp, -g

    def on_iterate(it, x, g, f):
        if it > 0:
            s, y = x - st["xprev"], g - st["gprev"]
            ys, yy = float(y @ s), float(y @ y)
            if ys > 0 and abs(yy / ys) <= 1e12:                                       # check_curve
                al = st["alpha"]
                yay, sias = float(y @ (al * y)), float(s @ (s / al))
                sa = s / al
                st["alpha"] = ys / (yay / al + y * y - (yay / sias) * sa * sa)       # form_diag
                st["S"].append(s)
                st["Y"].append(y)
                if len(st["S"]) > history:
                    st["S"].pop(0)
                    st["Y"].pop(0)
            ap = _build(st["alpha"], st["S"], st["Y"], x, g)
            elbo = -np.inf
            if ap is not None:
                U = np.stack([normal(seed, STREAM_PF_ELBO, (path << 24) | (it << 12) | k, idx) for k in range(num_elbo_draws)])
                phi, lq = _draw(ap, U)
                lp = ev.lp_batch(phi)
                # the oracle stops at the first non-finite draw; the mean over all draws is only used when every one is finite
                if np.all(np.isfinite(lp)) and np.all(np.isfinite(lq)):
                    elbo = float(np.sum(lp - lq)) / num_elbo_draws
            st["elbos"].append(elbo)
            if elbo > st["best_elbo"]:
                st["best_elbo"], st["best_iter"], st["best"] = elbo, it, ap
        st["xprev"], st["gprev"] = x.copy(), g.copy()

    x0 = init_radius * (2.0 * uniform(seed, STREAM_PF_INIT, path, idx) - 1.0)
    iters, term = lbfgs_path(fg, x0, max_lbfgs_iters, history, on_iterate)
    if st["best_iter"] < 0:
        raise FloatingPointError("Pathfinder: no iterate of the L-BFGS path produced a finite ELBO estimate")
    U = np.stack([normal(seed, STREAM_PF_DRAW, (path << 16) | k, idx) for k in range(num_draws)])
    phi, lq = _draw(st["best"], U)
    lp = ev.lp_batch(phi)
    lr = np.where(np.isfinite(lp), lp - lq, -np.inf)
    info = dict(lbfgs_iters=iters, lbfgs_term=term, best_iter=st["best_iter"], best_elbo=st["best_elbo"], n_pairs=st["best"].n,
                elbos=np.array(st["elbos"]))
    return phi, lr, info


def psis_weights(lr_in):
    """Pareto-smoothed importance weights [U: stan::services::psis::psis_weights; Vehtari et al. 2015, Zhang & Stephens 2009]"""
    lr = np.asarray(lr_in, dtype=np.float64) - np.max(lr_in)
    S = lr.shape[0]
    khat = np.nan
    tail = int(min(0.2 * S, 3.0 * np.sqrt(S)))
    if tail >= 5:
        order = np.argsort(lr, kind="stable")
        cutoff = lr[order[S - tail - 1]]
        ec = np.exp(cutoff)
        x = np.exp(lr[order[S - tail:]]) - ec
        n = tail
        if x[-1] > 0 and np.isfinite(x[-1]):
            M = 30 + int(np.floor(np.sqrt(n)))
            xstar, prior = x[int(np.floor(n / 4.0 + 0.5)) - 1], 3.0
            j = np.arange(M)
            with np.errstate(all="ignore"):
                th = 1.0 / x[-1] + (1.0 - np.sqrt(M / (j + 0.5))) / (prior * xstar)
                kk = np.mean(np.log1p(-th[:, None] * x[None, :]), axis=1)
                lt = n * (np.log(-th / kk) - kk - 1.0)
                wts = 1.0 / np.sum(np.exp(lt[None, :] - lt[:, None]), axis=1)
                theta = float(np.sum(th * wts))
                k = float(np.mean(np.log1p(-theta * x)))
                sigma = -k / theta
            k = k * n / (n + 10.0) + 0.5 * 10.0 / (n + 10.0)
            khat = k
            if np.isfinite(k) and np.isfinite(sigma):
                pq = (np.arange(n) + 0.5) / n
                q = (-sigma * np.log1p(-pq) if abs(k) < 1e-300 else sigma * np.expm1(-k * np.log1p(-pq)) / k) + ec
                lr[order[S - tail:]] = np.minimum(np.log(q), 0.0)
    w = np.exp(lr)
    return w / np.sum(w), khat


def gmm_pathfinder(X, K, n_draws=1000, elbo_samples=100, seed=312, num_paths=4, max_lbfgs_iters=500, history=100, init_radius=0.01,
                   device=0, evaluator=None, return_draws=True):
    """mod$pathfinder(...) on inst/clusterSVGs.stan + the posterior pass of R/clusterSVGsBayes.R:201-291.  Returns the dict
    engine.gmm_fit returns (mu / var are None: Pathfinder has no single variational family)."""
    import time
    t0 = time.time()
    ev = evaluator if evaluator is not None else DeviceEvaluator(X, K, device)
    try:
        phis, lrs, infos = [], [], []
        for pth in range(num_paths):
            phi, lr, info = single_path(ev, pth, seed, n_draws, max_lbfgs_iters, history, elbo_samples, init_radius)
            phis.append(phi)
            lrs.append(lr)
            infos.append(info)
        phi, lr = np.vstack(phis), np.concatenate(lrs)
        Stot = phi.shape[0]
        if num_paths > 1:
            w, khat = psis_weights(lr)
            uu = uniform(seed, STREAM_PF_PSIS, 0, np.arange(n_draws, dtype=np.uint64))
            cs = np.cumsum(w)
            idx = np.minimum(np.searchsorted(cs, uu, side="right"), Stot - 1)       # inverse-CDF resampling with replacement
        else:
            khat, idx = np.nan, np.arange(n_draws) % n_draws
        sel = phi[idx]
        draws, soft, ll = ev.posterior(sel, return_draws)
    finally:
        if evaluator is None:
            ev.close()
    info = dict(eta=float("nan"), iters=int(sum(i["lbfgs_iters"] for i in infos)), converged=1, elbo=float(max(i["best_elbo"] for i in infos)),
                log_likelihood=ll, seconds=time.time() - t0, grad_evals=0, elbo_evals=int(sum(len(i["elbos"]) for i in infos)),
                khat=float(khat), paths=infos, psis_index=idx)
    return dict(mu=None, var=None, draws=draws, soft=soft, info=info, unconstrained_draws=sel)

"""CPU: the oracle (oracle/bvg_oracle.c) against the committed golden vectors.

The golden vectors come from tests/golden/make_golden.py -- a literal, statement-by-statement
transcription of inst/approxGP.stan / inst/approxGP4.stan evaluated with scipy.stats + torch
autograd (fp64).  Tolerance: 1e-9 relative on log-density, 1e-8 norm-wise on gradients (fp64
round-off of two different summation orders).
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc

CASES = [("gp_small", "gaussian"), ("gp4_small", "nb"), ("gp_ragged", "gaussian"), ("gp4_ragged", "nb"),
         # depth-adjusted programs inst/approxGP2.stan / approxGP3.stan (gene.depth.adjust = TRUE)
         ("gp2_small", "gaussian"), ("gp3_small", "nb"), ("gp2_ragged", "gaussian"), ("gp3_ragged", "nb")]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name,model", CASES)
@pytest.mark.parametrize("nthreads", [1, 3])
def test_log_prob_and_grad_match_stan_transcription(golden_dir, name, model, nthreads):
    z = _load(golden_dir, name)
    o = orc.Oracle(z["phi"], z["y"], model, nthreads=nthreads, gene_depths=z["gene_depths"] if "gene_depths" in z.files else None)
    assert o.D == z["thetas"].shape[1]
    for jac in (0, 1):
        for pro in (0, 1):
            for t, th in enumerate(z["thetas"]):
                lp, g = o.log_prob(th, jacobian=jac, propto=pro)
                ref_lp, ref_g = z[f"lp_j{jac}_p{pro}"][t], z[f"grad_j{jac}_p{pro}"][t]
                assert abs(lp - ref_lp) <= 1e-9 * max(1.0, abs(ref_lp)), (name, jac, pro, t, lp, ref_lp)
                assert np.linalg.norm(g - ref_g) <= 1e-8 * np.linalg.norm(ref_g), (name, jac, pro, t)
                assert np.max(np.abs(g - ref_g) / (1e-6 + np.abs(ref_g))) < 1e-6
                lp2 = o.log_prob(th, jacobian=jac, propto=pro, grad=False)
                assert lp2 == pytest.approx(lp, rel=1e-13)


@pytest.mark.parametrize("name,model", CASES[:2] + CASES[4:6])
def test_gradient_central_differences(golden_dir, name, model):
    z = _load(golden_dir, name)
    o = orc.Oracle(z["phi"], z["y"], model, gene_depths=z["gene_depths"] if "gene_depths" in z.files else None)
    th = z["thetas"][1]
    _, g = o.log_prob(th, jacobian=1, propto=1)
    rng = np.random.default_rng(0)
    for i in rng.choice(o.D, 25, replace=False):
        h = 1e-5
        e = np.zeros(o.D)
        e[i] = h
        fd = (o.log_prob(th + e, grad=False) - o.log_prob(th - e, grad=False)) / (2 * h)
        assert fd == pytest.approx(g[i], rel=2e-5, abs=2e-5)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert orc.philox([0] * 4, [0] * 2) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_normal_moments():
    x = np.array([orc.normal(312, 1, 3, i) for i in range(100000)])
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1) < 0.01 and abs((x ** 4).mean() - 3) < 0.1
    assert orc.normal(312, 1, 3, 5) != orc.normal(312, 2, 3, 5) != orc.normal(313, 1, 3, 5)


def test_digamma():
    from scipy.special import digamma
    for x in (1e-3, 0.3, 1.0, 2.5, 9.99, 10.0, 57.0, 1e5):
        assert orc.digamma(x) == pytest.approx(digamma(x), rel=1e-13, abs=1e-13)


@pytest.mark.parametrize("kname,kernel,nu,period", [
    ("exp_quad", "exp_quad", 0.0, 1.0), ("matern0.5", "matern", 0.5, 1.0), ("matern1.5", "matern", 1.5, 1.0),
    ("matern2.5", "matern", 2.5, 1.0), ("periodic", "periodic", 0.0, 3.0)])
def test_basis_matches_lapack_dgeqp3(golden_dir, kname, kernel, nu, period):
    z = _load(golden_dir, "basis")
    Q, raw = orc.basis(z["coords"], z["centers"], float(z["lscale"]), kernel, nu, period, return_raw=True)
    # kernel values: closed forms == besselK form of R/maternKernel.R:22-26
    assert np.max(np.abs(raw - z[f"raw_{kname}"])) < 5e-15
    Qref = z[f"Q_{kname}"]
    assert np.max(np.abs(Q.T @ Q - np.eye(Q.shape[1]))) < 1e-12
    # same pivots + same Householder sign convention -> same Q, sign included
    cond = np.linalg.cond(z[f"raw_{kname}"])
    assert np.max(np.abs(Q - Qref)) < 1e-15 * cond * 50 + 1e-12, (kname, cond)


def test_summary_against_numpy():
    G, k, nd, seed = 7, 3, 1000, 312
    lay = orc.layout("gaussian", G, k)
    rng = np.random.default_rng(1)
    mu, om = rng.normal(size=lay["D"]), rng.normal(-1, 0.3, size=lay["D"])
    phi, y = np.linalg.qr(rng.normal(size=(11, k)))[0], rng.normal(size=(11, G))
    out = orc.Oracle(phi, y).summary(mu, om, nd, seed)
    for g in range(G):
        i = lay["amplitude"] + g
        x = np.exp(mu[i] + np.exp(om[i]) * np.array([orc.normal(seed, 3, t, i) for t in range(nd)]))
        med = np.median(x)
        ref = [x.mean(), med, x.std(ddof=1), 1.4826 * np.median(np.abs(x - med)), np.quantile(x, 0.05),
               np.quantile(x, 0.95), x.var(ddof=1) / x.mean()]
        assert np.allclose(out[g, :7], ref, rtol=1e-12)
    assert sorted(out[:, 7]) == list(range(1, G + 1))
    assert np.all(np.diff(out[np.argsort(out[:, 7]), 0]) <= 0)


def test_fullrank_family_restatement():
    """algorithm = "fullrank" [U: stan::variational::normal_fullrank]: with a diagonal L the family IS meanfield (same ELBO
    estimate from the same draws); the gradient of the reparameterised objective matches central differences; mu's step
    equals the meanfield step"""
    from tests.util import make_problem
    p = make_problem(60, 6, 3, "gaussian", seed=2)
    o = orc.Oracle(p["phi"], p["y"])
    D = o.D
    rng = np.random.default_rng(0)
    mu, om = rng.normal(0, 0.2, D), rng.normal(-2, 0.2, D)
    assert o.fr_elbo(mu, orc.pack_tril(np.diag(np.exp(om))), 20, 5, 3) == pytest.approx(o.elbo(mu, om, 20, 5, 3), rel=1e-13)
    Lf = np.tril(rng.normal(0, 0.05, (D, D)))
    Lf[np.diag_indices(D)] = np.exp(om)
    L0 = orc.pack_tril(Lf)
    assert np.array_equal(orc.unpack_tril(L0, D), Lf)
    eta = np.array([orc.normal(9, 1, 4, i) for i in range(D)])

    def obj(Lfull):   # log p(mu + L eta) + log |det L|: the one-draw ELBO estimate as a function of L
        return o.log_prob(mu + Lfull @ eta, True, True, grad=False) + np.log(np.abs(np.diag(Lfull))).sum()

    mu1, L1, hist = mu.copy(), L0.copy(), np.zeros(D + len(L0))
    assert o.fr_step(mu1, L1, hist, 1e-3, 1, 9, 4) == 0
    gabs = np.sqrt(hist)                      # iteration 1: the history IS the squared gradient
    ii, jj = np.tril_indices(D)
    for q in (0, 5, 100, 400, len(L0) - 1):
        Lp, Lm = Lf.copy(), Lf.copy()
        Lp[ii[q], jj[q]] += 1e-6
        Lm[ii[q], jj[q]] -= 1e-6
        fd = (obj(Lp) - obj(Lm)) / 2e-6
        assert abs(fd) == pytest.approx(gabs[D + q], rel=1e-6)
        assert np.sign(L1[q] - L0[q]) == np.sign(fd)          # ascent direction
    g = o.log_prob(mu + Lf @ eta, True, True)[1]
    assert np.allclose(mu1 - mu, 1e-3 * g / (1 + np.abs(g)), rtol=1e-12)
    # a short fit runs and its draws are finite
    m, L, info = o.advi_fullrank(np.zeros(D), iter=200, elbo_samples=20, seed=3)
    assert info["rc"] == 0 and np.isfinite(info["elbo"]) and np.all(np.isfinite(o.summary_fullrank(m, L, 50, 3)))


def test_pathfinder_restatement():
    """algorithm = "pathfinder" [U: stan pathfinder; Zhang et al. 2022]: the Gaussian approximation built from curvature
    pairs equals the BFGS recursion started at diag(alpha) (mean, covariance, log density); PSIS weights behave; a whole
    multi-path run produces finite summaries"""
    from scipy.stats import multivariate_normal
    from tests.util import make_problem
    rng = np.random.default_rng(0)
    D, n = 12, 4
    A = rng.normal(size=(D, D))
    Hs = A @ A.T + np.eye(D)
    S = rng.normal(size=(n, D))
    Y = S @ Hs                                          # exact curvature pairs of a quadratic: s'y > 0
    alpha, x, g = np.exp(rng.normal(0, 0.3, D)), rng.normal(size=D), rng.normal(size=D)
    u = np.vstack([np.eye(D), rng.normal(size=(3, D))])
    phi, lq, xc = orc.pf_approx(alpha, S, Y, x, g, u)
    H = np.diag(alpha)
    for s, y in zip(S, Y):                              # BFGS inverse-Hessian recursion
        rho = 1 / (y @ s)
        V = np.eye(D) - rho * np.outer(s, y)
        H = V @ H @ V.T + rho * np.outer(s, s)
    assert np.allclose(xc, x - H @ g, rtol=1e-12, atol=1e-13)
    Am = (phi[:D] - xc).T                               # phi - xc = A u with A A' = Sigma
    assert np.allclose(Am @ Am.T, H, rtol=1e-11, atol=1e-13)
    assert np.allclose(lq[D:], multivariate_normal(mean=xc, cov=H).logpdf(phi[D:]), rtol=1e-11)
    # n = 0: the approximation is N(x - alpha g, diag(alpha))
    phi0, lq0, xc0 = orc.pf_approx(alpha, np.zeros((0, D)), np.zeros((0, D)), x, g, u[D:])
    assert np.allclose(xc0, x - alpha * g) and np.allclose(phi0, xc0 + np.sqrt(alpha) * u[D:])
    assert np.allclose(lq0, multivariate_normal(mean=xc0, cov=np.diag(alpha)).logpdf(phi0), rtol=1e-12)
    # PSIS: weights sum to one; light-tailed ratios give a small k, heavy-tailed ones a large k
    w1, k1 = orc.psis_weights(rng.normal(0, 0.1, 2000))
    w2, k2 = orc.psis_weights(3.0 * rng.standard_t(2, 2000))
    assert w1.sum() == pytest.approx(1) and w2.sum() == pytest.approx(1) and k1 < 0.5 < k2 and w1.max() < 0.01
    p = make_problem(100, 8, 3, "gaussian", seed=1)
    o = orc.Oracle(p["phi"], p["y"])
    out, info = o.pathfinder(max_lbfgs_iters=40, history=6, num_elbo_draws=15, num_draws=100, num_paths=2, num_psis_draws=120, seed=2)
    assert info["rc"] == 0 and np.all(np.isfinite(out)) and all(q["best_iter"] >= 1 for q in info["paths"])
    assert orc.uniform(1, 5, 0, 0) != orc.uniform(1, 5, 1, 0) and 0 < orc.uniform(1, 5, 0, 0) < 1


def test_variogram_restatement():
    """lscale.estimator = "variogram" (:208-260): the empirical semivariogram against a brute-force double loop, and the fit
    recovers the range of data simulated from the model it fits"""
    rng = np.random.default_rng(3)
    M = 60
    X = rng.uniform(-1.7, 1.7, (M, 2))
    Z = rng.normal(size=(2, M))
    ls, ranges, tab = orc.variogram(X, Z, 1.5)
    cutoff = np.sqrt(((X.max(0) - X.min(0)) ** 2).sum()) / 3
    num, cnt, hs = np.zeros(15), np.zeros(15), np.zeros(15)
    for i in range(M):
        for j in range(i + 1, M):
            h = np.hypot(*(X[i] - X[j]))
            if 0 < h <= cutoff:
                b = min(int(h // (cutoff / 15)), 14)
                num[b] += (Z[0, i] - Z[0, j]) ** 2
                cnt[b] += 1
                hs[b] += h
    assert np.array_equal(cnt, tab["np"]) and np.allclose(hs / np.maximum(cnt, 1), tab["dist"])
    assert np.allclose(num / (2 * np.maximum(cnt, 1)), tab["gamma"][0])
    # data from a Matern-1.5 field with range 0.3 (gstat "Mat": rho(h / a)), plus a nugget
    M = 350
    X = rng.uniform(-1.7, 1.7, (M, 2))
    d = np.sqrt(((X[:, None] - X[None]) ** 2).sum(-1))
    K = (1 + d / 0.3) * np.exp(-d / 0.3) + 0.1 * np.eye(M)
    Z = (np.linalg.cholesky(K) @ rng.normal(size=(M, 40))).T
    ls, ranges, _ = orc.variogram(X, Z, 1.5)
    assert np.isfinite(ranges).sum() >= 35 and 0.15 < ls < 0.6


def test_gmm_restatement_matches_a_literal_transcription():
    """inst/clusterSVGs.stan statement by statement with scipy.stats densities (stick-breaking simplex transform with its
    Jacobian [U], lower = 0 transforms), and the analytic gradient against central differences"""
    from scipy.special import expit, logsumexp
    from scipy.stats import dirichlet, norm
    rng = np.random.default_rng(0)
    N, D, K = 60, 3, 4
    X = rng.normal(0, 2, (N, D))
    o = orc.GmmOracle(X, K)
    assert o.D == K - 1 + 2 * K * D

    def stan_lp(th):
        y, theta, st, lj = th[:K - 1], np.zeros(K), 1.0, 0.0
        for q in range(K - 1):
            z = expit(y[q] - np.log(K - 1 - q))
            theta[q] = st * z
            lj += np.log(st) + np.log(z) + np.log(1 - z)
            st -= theta[q]
        theta[K - 1] = st
        mu = th[K - 1:K - 1 + K * D].reshape(D, K).T                # matrix[K, D], column-major
        sig = np.exp(th[K - 1 + K * D:].reshape(K, D))              # array[K] vector[D]
        lj += th[K - 1 + K * D:].sum()
        lp = norm.logpdf(mu, 0, 7).sum() + norm.logpdf(sig, 0, 3).sum() + dirichlet.logpdf(theta, np.full(K, 2.0))
        for i in range(N):
            lp += logsumexp([np.log(theta[j]) + norm.logpdf(X[i], mu[j], sig[j]).sum() for j in range(K)])
        return lp, lj

    for scale in (0.0, 0.4, 1.2):
        th = rng.normal(0, 1, o.D) * scale
        lp, lj = stan_lp(th)
        assert o.log_prob(th, True, False, grad=False) == pytest.approx(lp + lj, rel=1e-12)
        assert o.log_prob(th, False, False, grad=False) == pytest.approx(lp, rel=1e-12)
        for jac in (False, True):
            g = o.log_prob(th, jac, True)[1]
            fd = np.array([(o.log_prob(th + 1e-6 * e, jac, True, grad=False) - o.log_prob(th - 1e-6 * e, jac, True, grad=False)) / 2e-6
                           for e in np.eye(o.D)])
            assert np.abs(fd - g).max() <= 1e-6 * max(1.0, np.abs(g).max())


@pytest.mark.parametrize("name", ["gmm_small", "gmm_default_shape"])
def test_gmm_oracle_matches_golden(golden_dir, name):
    """oracle (analytic gradient, hand-written reverse pass through the simplex transform) against the committed torch-autograd
    transcription of inst/clusterSVGs.stan (tests/golden/make_golden_gmm.py)"""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    o = orc.GmmOracle(z["X"], int(z["K"]))
    for jac in (0, 1):
        for pro in (0, 1):
            for t, th in enumerate(z["thetas"]):
                lp, g = o.log_prob(th, jac, pro)
                assert lp == pytest.approx(z[f"lp_j{jac}_p{pro}"][t], rel=1e-12), (jac, pro, t)
                assert np.linalg.norm(g - z[f"grad_j{jac}_p{pro}"][t]) <= 1e-11 * np.linalg.norm(g), (jac, pro, t)

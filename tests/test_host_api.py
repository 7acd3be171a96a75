"""CPU: the host-side mirror of the reference's R interface (argument checks, preprocessing, output
contract) and the gene-sharding helpers.  Messages are the reference's own
(R/findSpatiallyVariableFeaturesBayes.R:99-108)."""
import numpy as np
import pandas as pd
import pytest

import bayesvg_b200 as bvg
from bayesvg_b200 import dist, synth


def _obj(G=12, M=30, seed=0):
    rng = np.random.default_rng(seed)
    return bvg.SpatialData(rng.uniform(0, 100, (M, 2)), data=rng.normal(size=(G, M)), counts=rng.poisson(2.0, (G, M)))


def test_validation_messages_match_the_reference():
    o = _obj()
    f = bvg.findSpatiallyVariableFeaturesBayes
    with pytest.raises(ValueError, match="Please provide a spatial data object"):
        f(None, ["gene1"])
    with pytest.raises(ValueError, match="Please provide an object of class"):
        f(object(), ["gene1"])
    with pytest.raises(ValueError, match="Please identify a set of naive HVGs"):
        f(o, None)
    with pytest.raises(ValueError, match="Please provide a valid likelihood."):
        f(o, ["gene1"], likelihood="poisson")
    with pytest.raises(ValueError, match="Please provide a valid covariance kernel."):
        f(o, ["gene1"], kernel="rbf")
    with pytest.raises(ValueError, match="valid smoothness parameter"):
        f(o, ["gene1"], kernel="matern", kernel_smoothness=2.0)
    with pytest.raises(ValueError, match="valid variational inference approximation algorithm"):
        f(o, ["gene1"], algorithm="hmc")
    # accepted by the reference, rejected clearly here (no fallback)
    for kw in (dict(opencl_params=[0, 0]),):
        with pytest.raises(NotImplementedError):
            f(o, ["gene1"], **kw)
    with pytest.raises(KeyError):
        f(o, ["not_a_gene"], verbose=False)


def test_preprocessing_helpers():
    X = np.random.default_rng(1).uniform(0, 50, (200, 2))
    S = bvg.scale_coords(X)
    assert np.allclose(S.mean(0), 0) and np.allclose(S.std(0, ddof=1), 1)  # coop::scaler (:160)
    from oracle import oracle as orc
    C, info = orc.kmeans(S, 7, 100, 10, seed=3)    # the CPU restatement of bvg_kmeans (the product runs on the device only)
    assert C.shape == (7, 2) and np.all(np.abs(C) < 3) and info["converged"]
    lab = ((S[:, None, :] - C[None]) ** 2).sum(-1).argmin(1)
    assert np.allclose(C, np.array([S[lab == j].mean(0) for j in range(7)]), atol=1e-8)   # a Lloyd fixed point
    assert info["tot_withinss"] == pytest.approx(((S - C[lab]) ** 2).sum(), rel=1e-5)
    C1, i1 = orc.kmeans(S, 7, 100, 1, seed=3)
    assert i1["start"] == 0 and i1["tot_withinss"] >= info["tot_withinss"]               # nstart keeps the best start
    if bvg._lib.lib().bvg_device_count() == 0:
        with pytest.raises(bvg._lib.BvgError, match="no CUDA device"):                    # no CPU fallback
            bvg.kmeans_centers(S, 7, rng=3)
    d = np.sqrt(((C[:, None] - C[None]) ** 2).sum(-1))
    assert bvg.length_scale_kmeans(C) == pytest.approx(np.median(d[np.triu_indices(7, 1)]))  # :205-207
    with pytest.raises(ValueError):
        bvg.kmeans_centers(S[:3], 5)


def test_output_contract_consumers():
    """getBayesianGeneStats.R:60-68 / classifySVGs.R:60-82 on a table with the path's columns"""
    o = _obj(G=6)
    o.meta_features["amplitude_mean"] = [0.5, np.nan, 2.0, 0.1, np.nan, 1.0]
    o.meta_features["amplitude_mean_rank"] = [3, np.nan, 1, 4, np.nan, 2]
    st = bvg.getBayesianGeneStats(o)
    assert list(st["gene"]) == ["gene3", "gene6", "gene1", "gene4"]
    o = bvg.classifySVGs(o, n_SVG=2)
    assert o.variable_features == ["gene3", "gene6"]
    assert list(o.meta_features["svg_status"]) == [False, None, True, False, None, True]


def test_synthetic_shapes_match_baseline_configs():
    assert synth.visium_grid().shape == (4992, 2)  # BASELINE.json configs[1]
    assert synth.masked_blob(synth.visium_grid(), 2444).shape == (2444, 2)  # seu_brain, test_bayesVG.R:217
    X = bvg.scale_coords(synth.visium_grid(20, 16))
    C = synth.det_kmeans(X, 8)
    assert C.shape == (8, 2) and np.array_equal(C, synth.det_kmeans(X, 8))  # deterministic centroids
    phi = np.linalg.qr(np.random.default_rng(0).normal(size=(320, 8)))[0]
    Y, amp = synth.expression(phi, 40, "gaussian")
    assert Y.shape == (320, 40) and abs(Y.mean()) < 1e-12 and Y.std(ddof=1) == pytest.approx(1.0)
    Yc, _ = synth.expression(phi, 40, "nb")
    assert np.all(Yc >= 0) and np.all(Yc == np.floor(Yc)) and (amp == 0).mean() > 0.1


def test_gene_sharding_helpers():
    assert dist.gene_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert dist.gene_ranges(3000, 8)[-1] == (2625, 3000)
    for lik in ("gaussian", "nb"):
        G, k = 7, 3
        LG = dist.layout(lik, G, k)
        theta = np.arange(LG["D"], dtype=float)
        seen = np.zeros(LG["D"], dtype=int)
        for g0, g1 in dist.gene_ranges(G, 3):
            idx = dist.shard_index(lik, G, k, g0, g1)
            LL = dist.layout(lik, g1 - g0, k)
            assert idx.shape == (LL["D"],) and len(set(idx)) == LL["D"]
            loc = theta[idx]
            assert loc[LL["amplitude"]] == LG["amplitude"] + g0
            assert loc[LL["z_alpha_t"]] == LG["z_alpha_t"] + k * g0
            seen[idx] += 1
        sm = dist.shared_mask(lik, G, k)
        assert sm.sum() == 2 * k + 5 and np.all(seen[sm] == 3) and np.all(seen[~sm] == 1)


def test_depth_adjusted_layout_and_sharding_helpers():
    """inst/approxGP2.stan:13-23 / approxGP3.stan:13-23: beta0, beta1, tail, amplitude[G], mu_amplitude, sigma_amplitude,
    mu_alpha[k], sigma_alpha[k], z_alpha_t[k,G] -- no z_beta0, D = 5 + G + 2k + kG"""
    G, k = 9, 4
    for lik in ("gaussian", "nb"):
        L = dist.layout(lik, G, k, depth=True)
        assert L["D"] == 5 + G + 2 * k + k * G and "z_beta0" not in L
        assert (L["mu_beta0"], L["sigma_beta0"], L["tail"], L["amplitude"]) == (0, 1, 2, 3)
        assert L["z_alpha_t"] + k * G == L["D"]
        seen = np.zeros(L["D"], dtype=int)
        for g0, g1 in dist.gene_ranges(G, 2):
            idx = dist.shard_index(lik, G, k, g0, g1, depth=True)
            LL = dist.layout(lik, g1 - g0, k, depth=True)
            assert idx.shape == (LL["D"],) and len(set(idx)) == LL["D"]
            assert idx[LL["amplitude"]] == L["amplitude"] + g0 and idx[LL["z_alpha_t"]] == L["z_alpha_t"] + k * g0
            seen[idx] += 1
        sm = dist.shared_mask(lik, G, k, depth=True)
        assert sm.sum() == 2 * k + 5 and np.all(seen[sm] == 2) and np.all(seen[~sm] == 1)


def test_gene_depths_reach_the_fit_only_with_depth_adjust(monkeypatch):
    """gene.depth.adjust = TRUE: gene_depths = log(rowSums(counts[naive.hvgs, ])) (:293-303) is what the host passes down"""
    import bayesvg_b200.api as api
    o = _obj(G=6, M=40, seed=3)
    seen = {}

    class FakeFit:
        def __init__(self, phi, Y, lik, prec, eng, gene_depths=None, algorithm="meanfield", **cfg):
            seen["depths"], seen["shape"], seen["algorithm"] = gene_depths, Y.shape, algorithm

        def run(self):
            t = np.ones((seen["shape"][1], 8))
            t[:, 0] = np.arange(seen["shape"][1], 0, -1)
            return t

        def stats(self):
            return dict(eta=0.1, advi_iters=1, elbo=0.0)

        def close(self):
            pass

    monkeypatch.setattr(api.E, "Fit", FakeFit)
    monkeypatch.setattr(api.E, "kmeans", lambda X, k, *a: X[:k].copy())   # device entry points are faked: host logic only
    monkeypatch.setattr(api.E, "basis_build", lambda X, C, ls, *a: np.linalg.qr(np.random.default_rng(0).normal(size=(X.shape[0], C.shape[0])))[0])
    genes = ["gene2", "gene5", "gene1"]
    api.findSpatiallyVariableFeaturesBayes(o, genes, n_basis_fns=3, gene_depth_adjust=True, verbose=False)
    rows = [list(o.genes).index(g) for g in genes]
    assert np.allclose(seen["depths"], np.log(np.asarray(o.counts)[rows].sum(1)))
    api.findSpatiallyVariableFeaturesBayes(o, genes, n_basis_fns=3, verbose=False, algorithm="FullRank")
    assert seen["depths"] is None and seen["algorithm"] == "fullrank"       # :107 tolower(algorithm)


def test_variational_fit_object_mirrors_cmdstanvb(tmp_path):
    """save.model = TRUE / extractModel() (R/findSpatiallyVariableFeaturesBayes.R:483-489, R/extractModel.R:18-27):
    the host logic of the CmdStanVB stand-in on an oracle-made draws matrix (no device involved)."""
    from oracle import oracle as orc
    from tests.util import make_problem
    from bayesvg_b200.engine import Fit

    p = make_problem(40, 5, 3, "gaussian", seed=4)
    o = orc.Oracle(p["phi"], p["y"])
    rng = np.random.default_rng(1)
    mu, om = rng.normal(0, 0.3, o.D), rng.normal(-2, 0.2, o.D)
    P = o.posterior(mu, om, 50, seed=7)
    shell = Fit.__new__(Fit)                       # column names only need the shape and the model kind
    shell.G, shell.k, shell.D, shell._h = 5, 3, o.D, None
    shell.cfg = bvg.default_config(likelihood=0, depth_adjust=0)
    cols = shell.posterior_columns()
    assert cols[:5] == ["lp__", "log_p__", "log_g__", "mu_beta0", "sigma_beta0"] and cols[-1] == "amplitude_sq[5]"
    assert cols[3 + o.D - 1] == "sigma_y" and "z_alpha_t[2,4]" in cols and len(cols) == P.shape[1]
    L = orc.layout("gaussian", 5, 3)
    assert cols[3 + L["z_alpha_t"] + 3 * 3 + 1] == "z_alpha_t[2,4]"      # column-major: gene 4, basis fn 2
    vf = bvg.VariationalFit(cols, P, mu=mu, omega=om)
    assert vf.num_draws() == 50 and list(vf.draws("amplitude").columns) == [f"amplitude[{g}]" for g in range(1, 6)]
    assert np.array_equal(vf.lp(), P[:, 1]) and np.array_equal(vf.lp_approx(), P[:, 2])
    s = vf.summary("amplitude")
    ref = o.summary(mu, om, 50, seed=7)            # the reference's summary call, :419
    assert np.allclose(s[["mean", "median", "sd", "mad", "q5", "q95"]].to_numpy(), ref[:, :6], rtol=1e-12)
    a, a2 = vf.draws("amplitude[2]").to_numpy()[:, 0], vf.draws("amplitude_sq[2]").to_numpy()[:, 0]
    assert np.allclose(a * a, a2, rtol=1e-14)
    b0 = vf.draws(["mu_beta0", "sigma_beta0", "z_beta0[3]", "beta0[3]"]).to_numpy()
    assert np.allclose(b0[:, 0] + b0[:, 1] * b0[:, 2], b0[:, 3], rtol=1e-14)
    with pytest.raises(KeyError, match="Can't find"):
        vf.draws("nope")
    m = vf.mean_row()
    assert m[:3].tolist() == [0, 0, 0] and m[3 + L["sigma_y"]] == np.exp(mu[L["sigma_y"]]) and m[3] == mu[0]
    path = vf.save_csv(str(tmp_path / "vb.csv"))
    back = pd.read_csv(path, comment="#")
    assert "z_alpha_t.2.4" in back.columns and "amplitude_sq.5" in back.columns and len(back.columns) == len(cols)
    assert back.shape == (51, len(cols)) and np.allclose(back.to_numpy()[1:], P, rtol=1e-15)
    obj = _obj()
    assert bvg.extractModel(obj) is None
    obj.misc["model_fit"] = vf
    assert bvg.extractModel(obj) is vf
    with pytest.raises(ValueError, match="must be non-NULL"):
        bvg.extractModel(None)
    # the other three programs' headers (declaration order: approxGP4.stan:12-23, approxGP2/3.stan:13-23)
    for lik, depth, head in ((1, 0, ["mu_beta0", "sigma_beta0", "z_beta0[1]"]), (0, 1, ["beta0", "beta1", "sigma_y"]),
                             (1, 1, ["beta0", "beta1", "phi_nb"])):
        shell.cfg = bvg.default_config(likelihood=lik, depth_adjust=depth)
        shell.D = 5 + (1 if depth else 2) * 5 + 2 * 3 + 3 * 5
        c = shell.posterior_columns()
        assert c[3:6] == head and len(c) == 3 + shell.D + (1 if depth else 2) * 5
        Lx = orc.layout("nb" if lik else "gaussian", 5, 3, bool(depth))
        assert c[3 + Lx["amplitude"]] == "amplitude[1]" and c[3 + Lx["mu_alpha"]] == "mu_alpha[1]"
        assert c[3 + Lx["phi_nb" if lik else "sigma_y"]] == ("phi_nb" if lik else "sigma_y")


def test_cluster_svgs_bayes_host_logic(monkeypatch):
    """clusterSVGsBayes() (R/clusterSVGsBayes.R:63-137, :255-291): argument checks with the reference's messages, per-gene scaling +
    PCA, the cluster table, log-likelihood and BIC -- with the device fit faked (host logic only)"""
    import bayesvg_b200.api as api
    o = _obj(G=12, M=30, seed=4)
    f = bvg.clusterSVGsBayes
    with pytest.raises(ValueError, match="All arguments to clusterSVGsBayes"):
        f(o, None)
    with pytest.raises(ValueError, match="Please provide an object of class"):
        f(object(), ["gene1"])
    with pytest.raises(ValueError, match="Please provide a valid number of clusters."):
        f(o, ["gene1", "gene2"], n_clust=1)
    with pytest.raises(ValueError, match="valid variational inference approximation algorithm"):
        f(o, ["gene1", "gene2"], algorithm="hmc")
    with pytest.raises(NotImplementedError):
        f(o, ["gene1", "gene2"], opencl_params=[0, 0])
    seen = {}

    def fake_fit(X, K, algorithm, n_iter, n_draws, elbo_samples, seed, device=0):
        seen.update(X=X, K=K, algorithm=algorithm, n_iter=n_iter, n_draws=n_draws, elbo_samples=elbo_samples, seed=seed)
        N, D = X.shape
        soft = np.full((N, K), 0.1 / (K - 1))
        soft[np.arange(N), np.arange(N) % K] = 0.9
        return dict(mu=np.zeros(K - 1 + 2 * K * D), var=np.zeros(3), draws=np.ones((n_draws, K + 2 * K * D)), soft=soft,
                    info=dict(eta=0.1, iters=n_iter, converged=0, elbo=-1.0, log_likelihood=-50.0, seconds=0.0, grad_evals=0, elbo_evals=0))

    monkeypatch.setattr(api.E, "gmm_fit", fake_fit)
    svgs = [f"gene{i}" for i in (3, 1, 7, 9, 10, 12, 5)]
    res = f(o, svgs, n_clust=3, n_PCs=4, n_iter=500, n_draws=20, verbose=False)
    assert seen["K"] == 3 and seen["algorithm"] == "fullrank" and seen["elbo_samples"] == 300 and seen["seed"] == 312   # defaults (:67-76, :85-91)
    X = seen["X"]
    assert X.shape == (7, 4) and np.allclose(res["pca_embedding"], X)
    rows = [o.genes.index(g) for g in svgs]
    expr = np.asarray(o.data, dtype=np.float64)[rows]
    z = (expr - expr.mean(1, keepdims=True)) / expr.std(1, ddof=1, keepdims=True)          # t(coop::scaler(t(expr_mtx))) (:132)
    assert np.allclose(np.abs(X), np.abs(np.linalg.svd(z, full_matrices=False)[0][:, :4] * np.linalg.svd(z, compute_uv=False)[:4]))
    cdf = res["cluster_df"]
    assert list(cdf["gene"]) == svgs and list(cdf["assigned_cluster"]) == [1, 2, 3, 1, 2, 3, 1]       # which.max (:262-266), 1-based
    assert res["log_likelihood"] == -50.0 and res["BIC"] == pytest.approx(100.0 + (2 + 2 * 3 * 4) * np.log(7))   # :286-291
    fit = res["model_fit"]
    assert fit.columns[:3] == ["theta[1]", "theta[2]", "theta[3]"] and fit.columns[3] == "mu[1,1]" and fit.columns[-1] == "sigma[3,4]"
    # algorithm = "pathfinder" (:162-170): its own driver, elbo.samples defaults to 100 (:19-25), the reference's settings
    import bayesvg_b200.gmm_pathfinder as gpf
    seen2 = {}

    def fake_pf(X, K, **kw):
        seen2.update(kw, K=K)
        r = fake_fit(X, K, "pathfinder", 0, kw["n_draws"], kw["elbo_samples"], kw["seed"])
        r["info"].update(khat=0.3, paths=[{}] * kw["num_paths"])
        r["mu"] = r["var"] = None
        return r

    monkeypatch.setattr(gpf, "gmm_pathfinder", fake_pf)
    res = f(o, svgs, n_clust=3, n_PCs=4, n_draws=20, algorithm="pathfinder", verbose=False)
    assert (seen2["elbo_samples"], seen2["max_lbfgs_iters"], seen2["history"], seen2["init_radius"], seen2["num_paths"]) == (100, 500, 100, 0.01, 4)
    assert list(res["cluster_df"]["assigned_cluster"]) == [1, 2, 3, 1, 2, 3, 1]


def test_gmm_pathfinder_host_algorithm_matches_the_oracle():
    """bayesvg_b200/gmm_pathfinder.py (the host side of clusterSVGsBayes(algorithm = "pathfinder"): L-BFGS path, the Gaussian
    approximation of every iterate, ELBO selection, PSIS) against the oracle's C restatement of the same algorithm on
    inst/clusterSVGs.stan, with the ORACLE as the evaluator on both sides (the device evaluator is the GPU suite's part):
    same Philox streams -> same path, same ELBO estimates, same draws, same importance weights."""
    import bayesvg_b200.gmm_pathfinder as gpf
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    N, D, K = 150, 4, 3
    cent = rng.normal(0, 3, (K, D))
    X = np.vstack([cent[j] + rng.normal(0, 0.7, (N // K, D)) for j in range(K)])
    o = orc.GmmOracle(X, K)

    class Ev:
        P = o.D

        def lp_grad(self, th):
            return o.log_prob(th, True, True)

        def lp_batch(self, ths):
            return np.array([o.log_prob(t, True, True, grad=False) for t in ths])

    kw = dict(max_lbfgs_iters=30, history=6, num_elbo_draws=8, num_draws=40, seed=11)
    for path in (0, 1):
        phi_o, lr_o, io = o.pf_single_phi(path, **kw)
        phi_p, lr_p, ip = gpf.single_path(Ev(), path, 11, 40, 30, 6, 8, 0.01)
        assert io["rc"] == 0 and (io["lbfgs_iters"], io["best_iter"], io["n_pairs"]) == (ip["lbfgs_iters"], ip["best_iter"], ip["n_pairs"])
        eo = io["elbos"][1:1 + len(ip["elbos"])]
        assert np.allclose(eo, ip["elbos"], rtol=1e-9)
        assert np.linalg.norm(phi_o - phi_p) <= 1e-9 * np.linalg.norm(phi_o) and np.allclose(lr_o, lr_p, atol=1e-7)
    w_o, k_o = o.psis_weights(lr_o)
    w_p, k_p = gpf.psis_weights(lr_o)
    assert np.allclose(w_o, w_p, atol=1e-14) and k_o == pytest.approx(k_p, rel=1e-12)
    # the library's generator, restated in numpy for the host side
    assert gpf.uniform(312, 5, 7, np.arange(4, dtype=np.uint64)) == pytest.approx([orc.lib().bvgo_uniform(312, 5, 7, i) for i in range(4)], rel=1e-15)

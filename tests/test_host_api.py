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
    for kw in (dict(algorithm="fullrank"), dict(algorithm="pathfinder"), dict(opencl_params=[0, 0]),
               dict(lscale_estimator="variogram")):
        with pytest.raises(NotImplementedError):
            f(o, ["gene1"], **kw)
    with pytest.raises(KeyError):
        f(o, ["not_a_gene"], verbose=False)


def test_preprocessing_helpers():
    X = np.random.default_rng(1).uniform(0, 50, (200, 2))
    S = bvg.scale_coords(X)
    assert np.allclose(S.mean(0), 0) and np.allclose(S.std(0, ddof=1), 1)  # coop::scaler (:160)
    C = bvg.kmeans_centers(S, 7, rng=np.random.default_rng(3))
    assert C.shape == (7, 2) and np.all(np.abs(C) < 3)
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
        def __init__(self, phi, Y, lik, prec, eng, gene_depths=None, **cfg):
            seen["depths"], seen["shape"] = gene_depths, Y.shape

        def run(self):
            t = np.ones((seen["shape"][1], 8))
            t[:, 0] = np.arange(seen["shape"][1], 0, -1)
            return t

        def stats(self):
            return dict(eta=0.1, advi_iters=1, elbo=0.0)

        def close(self):
            pass

    monkeypatch.setattr(api.E, "Fit", FakeFit)
    monkeypatch.setattr(api.E, "basis_build", lambda X, C, ls, *a: np.linalg.qr(np.random.default_rng(0).normal(size=(X.shape[0], C.shape[0])))[0])
    genes = ["gene2", "gene5", "gene1"]
    api.findSpatiallyVariableFeaturesBayes(o, genes, n_basis_fns=3, gene_depth_adjust=True, verbose=False)
    rows = [list(o.genes).index(g) for g in genes]
    assert np.allclose(seen["depths"], np.log(np.asarray(o.counts)[rows].sum(1)))
    api.findSpatiallyVariableFeaturesBayes(o, genes, n_basis_fns=3, verbose=False)
    assert seen["depths"] is None

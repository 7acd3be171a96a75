"""GPU: the reference's own test script (tests/testthat/test_bayesVG.R:26-85, :155-185, :215-218), line for line where the
hot path is concerned, driven through the host mirror of the R interface -- a synthetic stand-in of seu_brain's shape
(2,444 spots; the .rda is absent from the reference mount) replaces the dataset.  Like the reference's tests these assert
classes and shapes; the numeric parity lives in test_gpu_parity.py."""
import numpy as np
import pandas as pd
import pytest

import bayesvg_b200 as bvg
from bayesvg_b200 import synth

pytestmark = pytest.mark.gpu

N_GENES, N_HVG, N_SVG = 900, 750, 300


@pytest.fixture(scope="module")
def brain():
    """stand-in for seu_brain: 2,444 in-tissue Visium spots, counts + log-normalised data (Seurat::NormalizeData)"""
    coords = synth.masked_blob(synth.visium_grid(), 2444)
    X = bvg.scale_coords(coords)
    C = synth.det_kmeans(X, 20, 312)
    phi = bvg.basis_build(X, C, bvg.length_scale_kmeans(C), "matern", 1.5)
    counts, _ = synth.expression(phi, N_GENES, "nb", 312)
    counts = counts.astype(np.int64).T                                   # genes x spots
    data = np.log1p(counts / np.maximum(counts.sum(0, keepdims=True), 1) * 1e4)
    genes = [f"Gene{i}" for i in range(N_GENES)]
    return coords, counts, data, genes


def _obj(brain):
    coords, counts, data, genes = brain
    return bvg.SpatialData(coords, data=data, counts=counts, genes=genes)


def test_kernels(brain):
    """test_bayesVG.R:37-56, :155-166: every kernel gives a double matrix with 20 columns and one row per spot"""
    coords = brain[0]
    spatial_mtx = bvg.scale_coords(coords)                                   # coop::scaler
    M, k = spatial_mtx.shape[0], 20
    kmeans_centers = bvg.kmeans_centers(spatial_mtx, k, 100, 10, rng=312)    # stats::kmeans(centers = k, iter.max = 100L, nstart = 10L)
    lscale = bvg.length_scale_kmeans(kmeans_centers)                         # median(dists_centers[upper.tri(dists_centers)])
    for kernel, kw in (("exp_quad", {}), ("matern", dict(nu=2.5)), ("periodic", dict(period=100.0))):
        phi = bvg.basis_build(spatial_mtx, kmeans_centers, lscale, kernel, **kw)   # qr.Q(qr(phi, LAPACK = TRUE))
        assert phi.dtype == np.float64 and phi.shape == (M, 20) and M == 2444
        if kernel != "periodic":   # period = 100 on scaled coordinates is numerically rank-deficient in the reference too
            assert np.allclose(phi.T @ phi, np.eye(k), atol=1e-10)


@pytest.mark.parametrize("likelihood", ["gaussian", "nb"])
def test_svg_model(brain, likelihood):
    """test_bayesVG.R:67-84, :169-185: fit, classify, read the table back, extract the saved fit"""
    obj = _obj(brain)
    genes = brain[3]
    naive_hvgs = list(np.random.default_rng(312).choice(genes, N_HVG, replace=False))     # getNaiveHVGs(n.hvg = 750L)
    obj = bvg.findSpatiallyVariableFeaturesBayes(obj, naive_hvgs=naive_hvgs, likelihood=likelihood, lscale_estimator="kmeans",
                                                 kernel="matern", kernel_smoothness=1.5, n_cores=1, save_model=True, verbose=False,
                                                 kmeans_rng=312)
    obj = bvg.classifySVGs(obj, n_SVG=N_SVG)
    svg_metadata = bvg.getBayesianGeneStats(obj)
    assert isinstance(svg_metadata, pd.DataFrame) and svg_metadata.shape[0] == N_HVG                 # expect_equal(nrow(...), 750)
    for col in ("gene", "amplitude_mean", "amplitude_median", "amplitude_sd", "amplitude_mad", "amplitude_ci_ll", "amplitude_ci_ul",
                "amplitude_dispersion", "amplitude_mean_rank", "svg_status"):
        assert col in svg_metadata.columns
    assert sorted(svg_metadata["amplitude_mean_rank"]) == list(range(1, N_HVG + 1))
    assert svg_metadata["amplitude_mean"].is_monotonic_decreasing and np.all(svg_metadata["amplitude_mean"] > 0)
    assert len(obj.variable_features) == N_SVG and int(svg_metadata["svg_status"].sum()) == N_SVG
    assert obj.meta_features.shape[0] == N_GENES and int(obj.meta_features["amplitude_mean"].isna().sum()) == N_GENES - N_HVG  # untested genes: NA
    svg_fit = bvg.extractModel(obj)                                                                     # expect_s3_class(svg_fit, "CmdStanVB")
    assert isinstance(svg_fit, bvg.VariationalFit) and svg_fit.num_draws() == 1000
    s = svg_fit.summary("amplitude")                                                                    # :419 fit_vi$summary(variables = "amplitude")
    tab = obj.meta_features.loc[naive_hvgs]
    assert np.allclose(s["mean"].to_numpy(), tab["amplitude_mean"].to_numpy(), rtol=1e-9)             # the table IS that summary
    assert np.allclose(s["q95"].to_numpy(), tab["amplitude_ci_ul"].to_numpy(), rtol=1e-9)
    d = svg_fit.draws(["amplitude[1]", "amplitude_sq[1]", "lp__", "log_p__", "log_g__"]).to_numpy()
    assert np.allclose(d[:, 0] ** 2, d[:, 1], rtol=1e-12) and np.all(d[:, 2] == 0) and np.all(np.isfinite(d[:, 3:]))


def test_dataset_dims(brain):
    """test_bayesVG.R:215-218 (ncol = 2,444 spots there)"""
    assert brain[1].shape == (N_GENES, 2444)


def test_other_algorithms_of_the_same_call(brain):
    """algorithm = "fullrank" / "pathfinder" (:108, :357-375) through the same host call, on a subset that keeps D small"""
    obj = _obj(brain)
    # full-rank ADVI moves D (D + 1) / 2 entries of L with one noisy gradient per step: as in CmdStan, eta adaptation only succeeds
    # for modest D ("All proposed step-sizes failed" otherwise), hence 8 genes x 5 basis functions (D = 76) for that algorithm
    for alg, hv, kb in (("fullrank", brain[3][:8], 5), ("pathfinder", brain[3][:40], 20)):
        o = bvg.findSpatiallyVariableFeaturesBayes(_obj(brain), naive_hvgs=hv, kernel="matern", kernel_smoothness=1.5, algorithm=alg,
                                                   n_basis_fns=kb, n_iter=300, n_draws=200, n_cores=2, verbose=False, kmeans_rng=312)
        st = bvg.getBayesianGeneStats(o)
        assert st.shape[0] == len(hv) and np.all(np.isfinite(st["amplitude_mean"])) and np.all(st["amplitude_mean"] > 0)
    hv = brain[3][:40]
    # test_bayesVG.R:59-66: the Seurat call of the reference's script estimates the length-scale from per-gene variograms
    o = bvg.findSpatiallyVariableFeaturesBayes(obj, naive_hvgs=hv, lscale_estimator="variogram", kernel="matern", kernel_smoothness=1.5,
                                               n_cores=1, n_iter=300, n_draws=200, verbose=False, kmeans_rng=312)
    assert bvg.getBayesianGeneStats(o).shape[0] == 40
    with pytest.raises(ValueError, match="length-scale estimator"):
        bvg.findSpatiallyVariableFeaturesBayes(obj, naive_hvgs=hv, lscale_estimator="idw", verbose=False)

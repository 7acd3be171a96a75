"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle and the
committed golden vectors.  Tolerances (BASELINE.json north_star): <= 1e-6 relative for the fp64
path, <= 1e-3 for the fp32 / TF32 fast path; what the kernels actually achieve is asserted tighter.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import make_problem, rel_err

pytestmark = pytest.mark.gpu

import bayesvg_b200 as bvg  # noqa: E402

TOL = {"fp64": 1e-6, "fast": 1e-3}
ACHIEVED = {"fp64": 1e-10, "fast": 2e-4}
GOLDEN = [("gp_small", "gaussian"), ("gp4_small", "nb"), ("gp_ragged", "gaussian"), ("gp4_ragged", "nb"),
          # gene.depth.adjust = TRUE: inst/approxGP2.stan / approxGP3.stan
          ("gp2_small", "gaussian"), ("gp3_small", "nb"), ("gp2_ragged", "gaussian"), ("gp3_ragged", "nb")]
ENGINES = [("fp64", "simt"), ("fast", "simt"), ("fast", "tcgen05")]


def _fit(phi, y, lik, prec, eng, **kw):
    try:
        return bvg.Fit(phi, y, lik, prec, eng, **kw)
    except bvg._lib.BvgError as e:
        if eng == "tcgen05" and "not built" in str(e):
            pytest.skip("tcgen05 engine not built")
        raise


@pytest.mark.parametrize("name,lik", GOLDEN)
@pytest.mark.parametrize("prec,eng", ENGINES)
def test_log_prob_golden(golden_dir, name, lik, prec, eng):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    f = _fit(z["phi"], z["y"], lik, prec, eng, gene_depths=z["gene_depths"] if "gene_depths" in z.files else None)
    for jac in (0, 1):
        for pro in (0, 1):
            for t, th in enumerate(z["thetas"]):
                lp, g = f.log_prob(th, jac, pro)
                ref_lp, ref_g = z[f"lp_j{jac}_p{pro}"][t], z[f"grad_j{jac}_p{pro}"][t]
                assert abs(lp - ref_lp) <= ACHIEVED[prec] * max(1.0, abs(ref_lp)), (jac, pro, t, lp, ref_lp)
                assert rel_err(g, ref_g) <= ACHIEVED[prec], (jac, pro, t)
                assert f.log_prob(th, jac, pro, grad=False) == pytest.approx(lp, rel=1e-12 if prec == "fp64" else 1e-5)
    f.close()


@pytest.mark.parametrize("lik", ["gaussian", "nb"])
@pytest.mark.parametrize("prec,eng", ENGINES)
@pytest.mark.parametrize("M,G,k", [(500, 300, 20), (97, 130, 7), (33, 1, 1), (1000, 257, 23)])
def test_log_prob_vs_oracle(lik, prec, eng, M, G, k):
    p = make_problem(M, G, k, lik, seed=M + G)
    o = orc.Oracle(p["phi"], p["y"], lik, nthreads=4)
    f = _fit(p["phi"], p["y"], lik, prec, eng)
    rng = np.random.default_rng(5)
    for scale in (0.0, 0.3, 1.0):
        th = rng.normal(0, 1, o.D) * scale
        lp, g = f.log_prob(th, True, True)
        rlp, rg = o.log_prob(th, True, True)
        assert abs(lp - rlp) <= TOL[prec] * max(1.0, abs(rlp))
        assert rel_err(g, rg) <= TOL[prec]
        assert abs(lp - rlp) <= ACHIEVED[prec] * max(1.0, abs(rlp)), (lp, rlp)
        assert rel_err(g, rg) <= ACHIEVED[prec] * 5
    f.close()


@pytest.mark.parametrize("lik", ["gaussian", "nb"])
@pytest.mark.parametrize("prec,eng", [("fp64", "simt"), ("fast", "simt"), ("fast", "auto")])
@pytest.mark.parametrize("M,G,k", [(300, 70, 31), (640, 45, 64), (700, 33, 200), (900, 40, 401), (1500, 300, 24)])
def test_log_prob_large_basis_vs_oracle(lik, prec, eng, M, G, k):
    """BASELINE config 5 has k ~ 400 basis functions: the general path (k_prep_big / k_gene_big with the CUDA-core
    kernel k_lik_gen or, fast + auto, the two tcgen05 GEMM kernels of k_lik_tcbig.cu)"""
    p = make_problem(M, G, k, lik, seed=M + k)
    o = orc.Oracle(p["phi"], p["y"], lik, nthreads=4)
    f = _fit(p["phi"], p["y"], lik, prec, eng)
    assert f.stats()["engine_used"] == (2 if eng == "auto" and prec == "fast" else 1)
    rng = np.random.default_rng(6)
    for scale, jac, pro in ((0.0, True, True), (0.2, True, False), (0.5, False, False)):
        th = rng.normal(0, 1, o.D) * scale
        lp, g = f.log_prob(th, jac, pro)
        rlp, rg = o.log_prob(th, jac, pro)
        assert abs(lp - rlp) <= ACHIEVED[prec] * max(1.0, abs(rlp)), (lp, rlp)
        assert rel_err(g, rg) <= ACHIEVED[prec] * 5
        assert f.log_prob(th, jac, pro, grad=False) == pytest.approx(lp, rel=1e-12 if prec == "fp64" else 1e-5)
    f.close()


def test_large_basis_advi_trajectory_and_fit():
    """same-Philox ADVI steps on the general path against the oracle (fp64), then a short full fit at k = 64"""
    p = make_problem(500, 24, 64, "gaussian", seed=77)
    o = orc.Oracle(p["phi"], p["y"], "gaussian", nthreads=4)
    f = _fit(p["phi"], p["y"], "gaussian", "fp64", "auto", seed=5)
    mu, om, hist = np.zeros(o.D), np.zeros(o.D), np.zeros(2 * o.D)
    f.set_variational(mu, None)
    f.advi_steps(0, 0.1, reset=True)
    for it in range(1, 13):
        o.advi_step(mu, om, hist, 0.1, it, 5, it - 1)
    f.advi_steps(12, 0.1)
    gmu, gom = f.get_variational()
    assert rel_err(gmu, mu) < 1e-9 and rel_err(gom, om) < 1e-9
    e, _ = f.elbo(8)
    assert np.isfinite(e)
    f.close()
    f = _fit(p["phi"], p["y"], "gaussian", "fast", "auto", seed=5, n_iter=300, mle_iter=40, n_draws=256, elbo_samples=20)
    tab = f.run()
    assert tab.shape == (24, 8) and np.all(np.isfinite(tab)) and np.all(tab[:, 0] > 0)
    f.close()


@pytest.mark.parametrize("lik", ["gaussian", "nb"])
@pytest.mark.parametrize("prec,eng", ENGINES)
@pytest.mark.parametrize("M,G,k", [(500, 300, 20), (97, 130, 7), (260, 50, 64)])
def test_depth_adjusted_models_vs_oracle(lik, prec, eng, M, G, k):
    """gene.depth.adjust = TRUE (inst/approxGP2.stan, approxGP3.stan): log density, gradient, same-Philox ADVI steps"""
    if eng == "tcgen05" and k > 23:
        pytest.skip("tcgen05 engine covers k <= 23")
    p = make_problem(M, G, k, lik, seed=M + G + 1)
    depths = np.log(np.random.default_rng(4).integers(300, 90000, G).astype(np.float64))
    o = orc.Oracle(p["phi"], p["y"], lik, nthreads=4, gene_depths=depths)
    f = _fit(p["phi"], p["y"], lik, prec, eng, gene_depths=depths, seed=9)
    assert f.D == o.D == 5 + G + 2 * k + k * G
    rng = np.random.default_rng(5)
    for scale, jac, pro in ((0.0, True, True), (0.3, True, False), (0.3, False, True), (1.0, False, False)):
        th = rng.normal(0, 1, o.D) * scale
        th[1] *= 0.05  # beta1 multiplies gene_depths ~ 6..11
        lp, g = f.log_prob(th, jac, pro)
        rlp, rg = o.log_prob(th, jac, pro)
        assert abs(lp - rlp) <= ACHIEVED[prec] * max(1.0, abs(rlp)), (lp, rlp)
        assert rel_err(g, rg) <= ACHIEVED[prec] * 5
    mu, om, hist = np.zeros(o.D), np.zeros(o.D), np.zeros(2 * o.D)
    f.set_variational(mu, None)
    f.advi_steps(0, 0.01, reset=True)
    for it in range(1, 9):
        o.advi_step(mu, om, hist, 0.01, it, 9, it - 1)
    f.advi_steps(8, 0.01)
    gmu, gom = f.get_variational()
    tol = 1e-9 if prec == "fp64" else 1e-3
    assert rel_err(gmu, mu) < tol and rel_err(gom, om) < tol
    e, _ = f.elbo(6)
    eo = o.elbo(gmu, gom, 6, 9, 0)  # same Philox stream / counters as the engine's first ELBO estimate
    assert e == pytest.approx(eo, rel=1e-10 if prec == "fp64" else 2e-4)
    f.close()


def test_depth_adjusted_full_fit_and_errors():
    p = make_problem(400, 60, 12, "gaussian", seed=8)
    depths = np.log(np.random.default_rng(1).integers(300, 90000, 60).astype(np.float64))
    with pytest.raises(bvg._lib.BvgError):
        bvg.Fit(p["phi"], p["y"], "gaussian", gene_depths=depths[:-1])     # wrong length
    with pytest.raises(bvg._lib.BvgError):
        bvg.Fit(p["phi"], p["y"], "gaussian", gene_depths=np.full(60, -np.inf))  # a gene with zero counts
    f = bvg.Fit(p["phi"], p["y"], "gaussian", "fast", "auto", gene_depths=depths, n_iter=400, mle_iter=60, n_draws=300,
                elbo_samples=30, seed=3)
    tab = f.run()
    assert tab.shape == (60, 8) and np.all(np.isfinite(tab[:, :7])) and np.all(tab[:, 0] > 0)
    with pytest.raises(bvg._lib.BvgError):
        f.set_y(p["y"][:, :10].copy(order="F"))  # the staged gene_depths no longer match
    f.close()


def test_nb_integer_input_equals_double_input():
    p = make_problem(200, 40, 5, "nb", seed=3)
    th = np.random.default_rng(0).normal(0, 0.3, 5 + 2 * 40 + 10 + 200)
    f1 = bvg.Fit(p["phi"], p["y"], "nb", "fp64", "simt")
    f2 = bvg.Fit(p["phi"], p["y"].astype(np.int32), "nb", "fp64", "simt")
    a, b = f1.log_prob(th), f2.log_prob(th)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("lik", ["gaussian", "nb"])
def test_advi_trajectory_matches_oracle_fp64(lik):
    """same Philox draws -> the first 30 stochastic-gradient steps agree to round-off"""
    p = make_problem(300, 50, 6, lik, seed=11)
    o = orc.Oracle(p["phi"], p["y"], lik, nthreads=4)
    f = bvg.Fit(p["phi"], p["y"], lik, "fp64", "simt", seed=312)
    rng = np.random.default_rng(1)
    mu0 = rng.normal(0, 0.1, o.D)
    mu, om, hist = mu0.copy(), np.zeros(o.D), np.zeros(2 * o.D)
    f.set_variational(mu0, None)
    f.advi_steps(0, 0.1, reset=True)
    for it in range(1, 31):
        o.advi_step(mu, om, hist, 0.1, it, 312, it - 1)
    f.advi_steps(30, 0.1, reset=False)
    gmu, gom = f.get_variational()
    assert rel_err(gmu, mu) < 1e-9 and rel_err(gom, om) < 1e-9
    e_gpu, _ = f.elbo(20)
    e_ref = o.elbo(mu, om, 20, 312, 0)
    assert e_gpu == pytest.approx(e_ref, rel=1e-9)
    f.close()


@pytest.mark.parametrize("prec,eng", ENGINES)
@pytest.mark.parametrize("M,G,k", [(300, 50, 6), (1000, 257, 20)])
def test_sufficient_statistic_elbo_equals_dense_elbo(prec, eng, M, G, k):
    """gaussian: the forward-only ELBO estimate from (sum y^2, Phi'^T y, Phi'^T Phi') must equal the dense
    pass over Y (same Philox draws) and the oracle's"""
    p = make_problem(M, G, k, "gaussian", seed=17)
    o = orc.Oracle(p["phi"], p["y"], "gaussian", nthreads=4)
    rng = np.random.default_rng(2)
    mu, om = rng.normal(0, 0.2, o.D), rng.normal(-2, 0.3, o.D)
    fd = _fit(p["phi"], p["y"], "gaussian", prec, eng, seed=5)
    fs = _fit(p["phi"], p["y"], "gaussian", prec, eng, seed=5, elbo_suffstat=1)
    for f in (fd, fs):
        f.set_variational(mu, om)
        f.advi_steps(0, 0.1, reset=True)
    ref = o.elbo(mu, om, 40, 5, 0)
    ed, es = fd.elbo(40)[0], fs.elbo(40)[0]
    tol = 1e-10 if prec == "fp64" else 2e-6
    assert es == pytest.approx(ref, rel=tol) and ed == pytest.approx(ref, rel=tol)
    # the draw counter advanced identically: a second estimate uses draws 40..79 on both paths
    assert fs.elbo(40)[0] == pytest.approx(fd.elbo(40)[0], rel=tol)
    assert fs.elbo(40)[0] == pytest.approx(o.elbo(mu, om, 40, 5, 80), rel=tol)
    fd.close()
    fs.close()


@pytest.mark.parametrize("prec,eng", ENGINES[1:])
def test_advi_trajectory_fast_path(prec, eng):
    p = make_problem(400, 150, 20, "gaussian", seed=12)
    o = orc.Oracle(p["phi"], p["y"], "gaussian", nthreads=4)
    f = _fit(p["phi"], p["y"], "gaussian", prec, eng, seed=7)
    mu, om, hist = np.zeros(o.D), np.zeros(o.D), np.zeros(2 * o.D)
    f.set_variational(mu, None)
    f.advi_steps(0, 0.1, reset=True)
    for it in range(1, 11):
        o.advi_step(mu, om, hist, 0.1, it, 7, it - 1)
    f.advi_steps(10, 0.1)
    gmu, gom = f.get_variational()
    assert rel_err(gmu, mu) < 1e-3 and rel_err(gom, om) < 1e-3
    f.close()


@pytest.mark.parametrize("lik", ["gaussian", "nb"])
@pytest.mark.parametrize("prec,eng", ENGINES)
def test_set_y_again_reuses_the_model_and_matches_a_fresh_fit(lik, prec, eng):
    """a second bvg_fit_set_y of the same shape keeps the allocations and restages only Y (double-buffered H2D)"""
    p1 = make_problem(333, 140, 9, lik, seed=31)
    p2 = make_problem(333, 140, 9, lik, seed=32)
    f = _fit(p1["phi"], p1["y"], lik, prec, eng)
    th = np.random.default_rng(2).normal(0, 0.3, f.D)
    f.log_prob(th)
    f.set_y(p2["y"])                       # same shape: reuse path
    lp_a, g_a = f.log_prob(th, True, False)
    f.set_y(p2["y"][:, :77].copy(order="F"))  # different G: full re-allocation
    th77 = np.random.default_rng(3).normal(0, 0.3, f.D)
    lp_c, g_c = f.log_prob(th77, True, False)
    f.close()
    f2 = _fit(p1["phi"], p2["y"], lik, prec, eng)
    lp_b, g_b = f2.log_prob(th, True, False)
    f2.close()
    f3 = _fit(p1["phi"], p2["y"][:, :77].copy(order="F"), lik, prec, eng)
    lp_d, g_d = f3.log_prob(th77, True, False)
    f3.close()
    assert lp_a == lp_b and np.array_equal(g_a, g_b)
    assert lp_c == lp_d and np.array_equal(g_c, g_d)


def test_recycled_device_blocks_do_not_leak_state_between_fits():
    """bvg_fit_destroy keeps the large device blocks (staged Y, staging buffers, L-BFGS work space) for the next fit of the same
    sizes.  A fit that receives recycled blocks -- after a fit on DIFFERENT data of the same shape, whose padded entries and
    history rows are still in them -- must give bit-identical results to a fit on fresh memory; bvg_release_cached_memory empties
    the cache."""
    from bayesvg_b200 import _lib
    kw = dict(seed=312, n_iter=200, mle_iter=40, elbo_samples=20, n_draws=100)
    pa = make_problem(700, 300, 20, "gaussian", seed=1)        # 300 genes in 3 tiles of 128: 84 padded gene rows
    pb = make_problem(700, 300, 20, "gaussian", seed=2)
    _lib.lib().bvg_release_cached_memory()
    f = bvg.Fit(pa["phi"], pa["y"], "gaussian", "fast", "tcgen05", **kw)
    ref = f.run()
    f.close()
    f = bvg.Fit(pb["phi"], pb["y"], "gaussian", "fast", "tcgen05", **kw)               # other data into the same blocks
    f.run()
    f.close()
    f = bvg.Fit(pa["phi"], pa["y"], "gaussian", "fast", "tcgen05", **kw)               # recycled blocks
    got = f.run()
    f.close()
    assert np.array_equal(ref, got)
    _lib.lib().bvg_release_cached_memory()
    f = bvg.Fit(pa["phi"], pa["y"], "gaussian", "fast", "tcgen05", **kw)               # fresh memory again
    assert np.array_equal(ref, f.run())
    f.close()


def test_default_elbo_path_is_the_sufficient_statistic_one(monkeypatch):
    """bvg_config_default: elbo_suffstat = 1 -- a default gaussian fit takes its ELBO estimates from the sufficient statistics
    and ends where the dense-ELBO fit ends (same draws, values equal to the fast path's round-off: same step size, same stopping
    iteration); nb and depth-adjusted fits take the dense kernel whatever the flag says."""
    monkeypatch.delenv("BVG_ELBO_SUFFSTAT", raising=False)
    assert bvg.engine.default_config().elbo_suffstat == 1
    p = make_problem(500, 150, 12, "gaussian", seed=7)
    kw = dict(seed=312, n_iter=400, mle_iter=100, elbo_samples=30, n_draws=100)
    fa = bvg.Fit(p["phi"], p["y"], "gaussian", "fast", "tcgen05", **kw)
    fb = bvg.Fit(p["phi"], p["y"], "gaussian", "fast", "tcgen05", elbo_suffstat=0, **kw)
    ta, tb = fa.run(), fb.run()
    sa, sb = fa.stats(), fb.stats()
    assert (sa["eta"], sa["advi_iters"], sa["converged"]) == (sb["eta"], sb["advi_iters"], sb["converged"])
    assert sa["elbo"] == pytest.approx(sb["elbo"], rel=5e-6) and sa["elbo"] != sb["elbo"]     # fp64 statistics vs the bf16-split pass
    assert np.allclose(ta[:, :6], tb[:, :6], rtol=1e-9, atol=0)     # the ELBO estimates only steer eta and the stopping test
    fa.close()
    fb.close()
    pn = make_problem(300, 140, 12, "nb", seed=9)
    fn = bvg.Fit(pn["phi"], pn["y"], "nb", "fast", "auto", seed=3, n_iter=100, mle_iter=10, elbo_samples=10, n_draws=50)
    assert np.all(np.isfinite(fn.run()[:, :6]))
    fn.close()


@pytest.mark.parametrize("lik", ["gaussian", "nb"])
def test_draw_ahead_batches_are_bitwise_identical(lik):
    """tcgen05 engine: inside a batch of steps k_gene draws the next evaluation's zeta itself; a batch boundary
    re-draws it with k_prep.  Both must give the same numbers: 10 steps == 3 + 1 + 6 steps, bit for bit, and an
    ELBO estimate in between (which overwrites zeta with its own draws) must not disturb the trajectory."""
    p = make_problem(300, 200, 17, lik, seed=21)
    out = []
    for batches in ((10,), (3, 1, 6)):
        f = _fit(p["phi"], p["y"], lik, "fast", "tcgen05", seed=11)
        f.set_variational(np.zeros(f.D), None)
        f.advi_steps(0, 0.05, reset=True)
        for i, n in enumerate(batches):
            f.advi_steps(n, 0.05)
            if i == 0 and len(batches) > 1:
                f.elbo(3)
        out.append(f.get_variational())
        f.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def _advi_run(p, lik, steps, no_persist, depths=None, seed=11, eta=0.05):
    """(mu, omega, kernel launches) after `steps` mean-field iterations on the fused tcgen05 engine"""
    old = os.environ.get("BVG_NO_PERSIST")
    os.environ["BVG_NO_PERSIST"] = "1" if no_persist else "0"
    try:
        f = _fit(p["phi"], p["y"], lik, "fast", "tcgen05", seed=seed, gene_depths=depths)
    finally:
        if old is None:
            del os.environ["BVG_NO_PERSIST"]
        else:
            os.environ["BVG_NO_PERSIST"] = old
    f.set_variational(np.zeros(f.D), None)
    f.advi_steps(0, eta, reset=True)
    n0 = f.stats()["kernel_launches"]
    for n in steps:
        f.advi_steps(n, eta)
    out = f.get_variational() + (f.stats()["kernel_launches"] - n0,)
    f.close()
    return out


@pytest.mark.parametrize("lik,depth", [("gaussian", False), ("nb", False), ("gaussian", True), ("nb", True)])
@pytest.mark.parametrize("M,G,k", [(1000, 257, 20), (300, 200, 17), (97, 130, 7), (260, 1, 1), (4000, 700, 23), (9000, 300, 12)])
def test_persistent_advi_kernel_equals_the_two_launch_chain(lik, depth, M, G, k):
    """k_advi_persist (one cooperative launch per stretch: passes, tile / grid arrival counters, replicated finish) against
    k_lik_tc -> k_gene per evaluation: same Philox draws, same per-gene arithmetic; only the cross-gene sums are added in a
    different fixed order, so 25 iterations agree to rounding.  Also: the stretch really is 2 launches (k_prep + persistent).
    (9000, 300, 12): 3 tiles x 47 spot splits -- more splits than a tile has genes to hand out, so some CTAs own no genes."""
    p = make_problem(M, G, k, lik, seed=21)
    depths = np.log(np.maximum(p["y"].sum(0), 1.0)) - 5.0 if depth else None
    mu_a, om_a, la = _advi_run(p, lik, (25,), False, depths)
    mu_b, om_b, lb = _advi_run(p, lik, (25,), True, depths)
    assert la == 2 and lb >= 50
    assert np.all(np.isfinite(mu_a)) and rel_err(mu_a, mu_b) < 1e-9 and rel_err(om_a, om_b) < 1e-9
    # stretch boundaries re-draw with k_prep: 25 == 7 + 1 + 17, bit for bit
    mu_c, om_c, _ = _advi_run(p, lik, (7, 1, 17), False, depths)
    assert np.array_equal(mu_a, mu_c) and np.array_equal(om_a, om_c)


@pytest.mark.parametrize("lik", ["gaussian", "nb"])
def test_persistent_advi_kernel_row_form_of_the_grid_meeting(lik):
    """The cross-gene sums normally meet in fixed-point accumulators; sums outside their range (|s| >= 2^84) or not finite take
    the row form (every CTA publishes its row, second meeting).  BVG_PERSIST_FORCE_ROWS=1 takes it on every iteration: same
    trajectory as the two-launch chain and as the accumulator form."""
    p = make_problem(1000, 257, 20, lik, seed=21)
    mu_a, om_a, _ = _advi_run(p, lik, (7, 1, 17), False)
    os.environ["BVG_PERSIST_FORCE_ROWS"] = "1"
    try:
        mu_b, om_b, _ = _advi_run(p, lik, (7, 1, 17), False)
    finally:
        del os.environ["BVG_PERSIST_FORCE_ROWS"]
    mu_c, om_c, _ = _advi_run(p, lik, (25,), True)
    assert np.all(np.isfinite(mu_b))
    assert rel_err(mu_b, mu_c) < 1e-9 and rel_err(om_b, om_c) < 1e-9
    assert rel_err(mu_a, mu_b) < 1e-12 and rel_err(om_a, om_b) < 1e-12


def test_persistent_advi_kernel_survives_the_diverging_trials_of_adapt_eta():
    """adapt_eta's first trials (eta = 100, 10) drive a few genes' amplitudes to 1e4 and beyond: cross-gene sums of 1e13 ... inf.
    The accumulator form of the grid meeting must hand those iterations to the row form instead of poisoning the fit (a
    192-gene fit on 2,444 spots failed with 'non-finite gradient' when the accumulators covered |s| < 2^42 only and turned
    everything above into NaN): the persistent kernel and the two-launch chain end at the same ELBO."""
    from bayesvg_b200 import synth
    from bayesvg_b200.api import kmeans_centers, length_scale_kmeans
    X = synth.scale_coords(synth.masked_blob(synth.visium_grid(), 2444))
    Cc = kmeans_centers(X, 20, 100, 10, rng=312, device=0)
    phi = bvg.basis_build(X, Cc, length_scale_kmeans(Cc), "matern", 2.5)
    Y, _ = synth.expression(phi, 192, "gaussian", 312)
    out = {}
    for nop in ("0", "1"):
        os.environ["BVG_NO_PERSIST"] = nop
        try:
            f = bvg.Fit(phi, np.asfortranarray(Y), "gaussian", "fast", "auto", seed=312)
            f.run()
            out[nop] = f.stats()
            f.close()
        finally:
            del os.environ["BVG_NO_PERSIST"]
    assert out["0"]["advi_iters"] == out["1"]["advi_iters"] and out["0"]["eta"] == out["1"]["eta"]
    assert out["0"]["elbo"] == pytest.approx(out["1"]["elbo"], rel=1e-6)


@pytest.mark.parametrize("lik", ["gaussian", "nb"])
def test_persistent_advi_kernel_with_more_work_items_than_sms(lik):
    """157 gene tiles (x spot splits) on 148 SMs: a CTA of the persistent kernel walks several (tile, split) items per iteration"""
    p = make_problem(700, 20000, 5, lik, seed=3)
    mu_a, om_a, la = _advi_run(p, lik, (12,), False)
    mu_b, om_b, _ = _advi_run(p, lik, (12,), True)
    assert la == 2
    assert np.all(np.isfinite(mu_a)) and rel_err(mu_a, mu_b) < 1e-9 and rel_err(om_a, om_b) < 1e-9


@pytest.mark.parametrize("iters,tol_lp,tol_th", [(8, 1e-10, 1e-7), (60, 1e-5, 5e-2)])
def test_lbfgs_matches_oracle_fp64(iters, tol_lp, tol_th):
    """same algorithm (two-loop vs its vector-free Gram form): identical iterates early on, the same
    optimum later (round-off is amplified along the line searches)"""
    p = make_problem(300, 40, 5, "gaussian", seed=21)
    o = orc.Oracle(p["phi"], p["y"], "gaussian", nthreads=4)
    f = bvg.Fit(p["phi"], p["y"], "gaussian", "fp64", "simt", mle_iter=iters)
    th = f.mle()
    ref, info = o.lbfgs(max_iter=iters)
    st = f.stats()
    assert st["mle_iters"] == info["iters"]
    assert st["mle_lp"] == pytest.approx(info["lp"], rel=tol_lp)
    assert rel_err(th, ref) < tol_th
    lp0 = o.log_prob(np.zeros(o.D), False, False, grad=False)
    assert st["mle_lp"] > lp0
    assert o.log_prob(th, False, False, grad=False) == pytest.approx(st["mle_lp"], rel=1e-10)
    f.close()


@pytest.mark.parametrize("nd", [1000, 257])
def test_summary_matches_oracle(nd):
    p = make_problem(64, 33, 4, "gaussian", seed=2)
    o = orc.Oracle(p["phi"], p["y"])
    rng = np.random.default_rng(3)
    mu, om = rng.normal(0, 1, o.D), rng.normal(-1, 0.5, o.D)
    f = bvg.Fit(p["phi"], p["y"], "gaussian", "fp64", "simt", n_draws=nd, seed=99)
    f.set_variational(mu, om)
    got = f.summarize()
    ref = o.summary(mu, om, nd, 99)
    assert np.allclose(got[:, :7], ref[:, :7], rtol=1e-10, atol=0)
    assert np.array_equal(got[:, 7], ref[:, 7])
    d = f.draws()
    assert d.shape == (nd, 33) and np.allclose(d.mean(0), ref[:, 0], rtol=1e-10)
    f.close()


@pytest.mark.parametrize("lik,depth", [("gaussian", False), ("nb", False), ("gaussian", True), ("nb", True)])
@pytest.mark.parametrize("prec,eng", ENGINES)
def test_posterior_draws_of_the_whole_model(lik, depth, prec, eng):
    """save.model = TRUE (R/findSpatiallyVariableFeaturesBayes.R:483-489): the draws matrix behind the CmdStanVB object --
    lp__, log_p__, log_g__, parameters constrained in declaration order, beta0[G], amplitude_sq[G]."""
    p = make_problem(97, 21, 6, lik, seed=11)
    gd = np.log(np.abs(p["y"]).sum(0) + 1.0) if depth else None
    o = orc.Oracle(p["phi"], p["y"], lik, gene_depths=gd)
    rng = np.random.default_rng(8)
    mu, om = rng.normal(0, 0.3, o.D), rng.normal(-2.5, 0.3, o.D)
    f = _fit(p["phi"], p["y"], lik, prec, eng, gene_depths=gd, n_draws=12, seed=41)
    f.set_variational(mu, om)
    cols = f.posterior_columns()
    got, ref = f.posterior(12), o.posterior(mu, om, 12, seed=41)
    assert got.shape == ref.shape == (12, len(cols))
    assert np.all(got[:, 0] == 0)
    assert np.allclose(got[:, 2:], ref[:, 2:], rtol=1e-11, atol=1e-13)                    # fp64 state in both precisions
    assert np.allclose(got[:, 1], ref[:, 1], rtol=ACHIEVED[prec], atol=0)                 # log_p__ goes through the engine
    f.summarize()
    ia = cols.index("amplitude[1]")
    assert np.array_equal(got[:, ia:ia + 21], f.draws())                                  # same Philox stream as the summary
    assert np.array_equal(f.posterior(5), got[:5])                                        # repeatable, counter restarts
    lp0, _ = f.log_prob(mu, True, False)                                                  # the handle is still usable
    assert np.isfinite(lp0)
    f.close()


@pytest.mark.parametrize("kname,kernel,nu,period", [
    ("exp_quad", "exp_quad", 0.0, 1.0), ("matern0.5", "matern", 0.5, 1.0), ("matern1.5", "matern", 1.5, 1.0),
    ("matern2.5", "matern", 2.5, 1.0), ("periodic", "periodic", 0.0, 3.0)])
def test_basis_matches_lapack_golden(golden_dir, kname, kernel, nu, period):
    z = np.load(os.path.join(golden_dir, "basis.npz"))
    Q = bvg.basis_build(z["coords"], z["centers"], float(z["lscale"]), kernel, nu if kernel == "matern" else 1.5, period)
    cond = np.linalg.cond(z[f"raw_{kname}"])
    assert np.max(np.abs(Q.T @ Q - np.eye(Q.shape[1]))) < 1e-12
    assert np.max(np.abs(Q - z[f"Q_{kname}"])) < 1e-15 * cond * 50 + 1e-12


@pytest.mark.parametrize("M,k,nstart,iter_max", [(2444, 20, 10, 100), (700, 7, 3, 4), (50, 50, 1, 10), (20000, 400, 2, 100), (33, 1, 2, 5)])
def test_kmeans_centroids_equal_the_cpu_restatement(M, k, nstart, iter_max):
    """stats::kmeans(spatial_mtx, centers = k, iter.max = 100, nstart = 10) (R/findSpatiallyVariableFeaturesBayes.R:200-204) on
    the device: fixed-point sums make it bit-reproducible, so the centroids equal the serial restatement exactly."""
    X = bvg.scale_coords(np.random.default_rng(M).uniform(0, 100, (M, 2)))
    C, info = bvg.engine.kmeans(X, k, iter_max, nstart, seed=M + 1, return_info=True)
    ref, rinfo = orc.kmeans(X, k, iter_max, nstart, seed=M + 1)
    assert np.array_equal(C, ref) and info == rinfo
    assert np.array_equal(C, bvg.engine.kmeans(X, k, iter_max, nstart, seed=M + 1))       # run-to-run identical
    lab = ((X[:, None, :] - C[None]) ** 2).sum(-1).argmin(1) if M * k < 2e6 else None
    if lab is not None and info["converged"]:
        assert np.allclose(C, np.array([X[lab == j].mean(0) if np.any(lab == j) else C[j] for j in range(k)]), atol=1e-8)
        assert info["tot_withinss"] == pytest.approx(((X - C[lab]) ** 2).sum(), rel=1e-4)
    assert bvg.kmeans_centers(X, k, rng=5).shape == (k, 2)
    with pytest.raises(ValueError, match="more cluster centers"):
        bvg.kmeans_centers(X[:3], 5)
    with pytest.raises(bvg._lib.BvgError, match="scaled"):
        bvg.engine.kmeans(X * 1e4, k)


@pytest.mark.parametrize("kappa", [0.5, 1.5, 2.5])
def test_variogram_length_scale_matches_the_cpu_restatement(kappa):
    """lscale.estimator = "variogram" (R/findSpatiallyVariableFeaturesBayes.R:208-260): per-gene empirical semivariograms and
    Matern variogram fits on the device against the numpy / scipy restatement (same classes, same weights, same start)"""
    rng = np.random.default_rng(7)
    M, G = 300, 37
    X = bvg.scale_coords(rng.uniform(0, 100, (M, 2)))
    d = np.sqrt(((X[:, None] - X[None]) ** 2).sum(-1))
    K = (1 + d / 0.25) * np.exp(-d / 0.25) + 0.15 * np.eye(M)
    Z = (np.linalg.cholesky(K) @ rng.normal(size=(M, G))).T * rng.uniform(0.5, 2, G)[:, None] + rng.normal(size=(G, 1))
    Z[5] = 3.0                                                  # a constant gene: no variogram, NA in the reference
    ls, ranges = bvg.engine.variogram_lscale(X, Z, kappa, return_ranges=True)
    rls, rranges, _ = orc.variogram(X, Z, kappa)
    assert np.isnan(ranges[5]) and np.isnan(rranges[5])
    both = np.isfinite(ranges) & np.isfinite(rranges)
    assert both.sum() >= G - 4
    close = np.isclose(ranges[both], rranges[both], rtol=1e-4)
    assert close.mean() >= 0.9, (ranges[both][~close], rranges[both][~close])    # the two optimisers may part on a flat objective
    assert ls == pytest.approx(rls, rel=1e-3)
    assert ls == bvg.engine.variogram_lscale(X, Z, kappa)                        # deterministic
    with pytest.raises(bvg._lib.BvgError, match="smoothness"):
        bvg.engine.variogram_lscale(X, Z, 2.0)


def test_basis_visium_shape():
    from bayesvg_b200 import synth
    X = bvg.scale_coords(synth.visium_grid())
    C = synth.det_kmeans(X, 20)
    Q = bvg.basis_build(X, C, bvg.length_scale_kmeans(C), "matern", 2.5)
    assert Q.shape == (4992, 20) and Q.dtype == np.float64  # tests/testthat/test_bayesVG.R:156-166
    ref = orc.basis(X, C, bvg.length_scale_kmeans(C), "matern", 2.5)
    assert np.max(np.abs(Q - ref)) < 1e-10


@pytest.mark.parametrize("prec,eng", ENGINES)
def test_full_fit_agrees_with_cpu_restatement(prec, eng):
    """whole path (L-BFGS init -> adapt -> SGA -> draws -> summary) vs the oracle driving the same
    algorithm with the same Philox streams: the fp64 path must reproduce the oracle's decisions;
    the fast path must agree within Monte Carlo tolerance (Spearman, top-k overlap)."""
    from scipy.stats import spearmanr
    p = make_problem(500, 150, 12, "gaussian", seed=7)
    o = orc.Oracle(p["phi"], p["y"], "gaussian", nthreads=4)
    cfg = dict(n_iter=400, mle_iter=100, elbo_samples=30, n_draws=500, seed=312)
    f = _fit(p["phi"], p["y"], "gaussian", prec, eng, **cfg)
    got = f.run()
    st = f.stats()
    th, li = o.lbfgs(max_iter=100)
    mu, om, info = o.advi(th, iter=400, elbo_samples=30, seed=312)
    ref = o.summary(mu, om, 500, 312)
    assert st["eta"] == info["eta"]
    rho = spearmanr(got[:, 0], ref[:, 0]).correlation
    top = 40
    ov = len(set(np.argsort(-got[:, 0])[:top]) & set(np.argsort(-ref[:, 0])[:top])) / top
    if prec == "fp64":
        assert st["advi_iters"] == info["iters"] and st["converged"] == info["converged"]
        assert rho > 0.999 and ov >= 0.95
    else:
        assert rho > 0.98 and ov >= 0.9, (rho, ov)
    tr = f.trace()
    assert tr.shape[1] == 4 and len(tr) == st["advi_iters"] // 100
    f.close()


@pytest.mark.parametrize("lik", ["gaussian", "nb"])
@pytest.mark.parametrize("prec,eng", ENGINES)
def test_fullrank_trajectory_elbo_and_draws_match_oracle(lik, prec, eng):
    """algorithm = "fullrank" (R/findSpatiallyVariableFeaturesBayes.R:108, :357-364): q = N(mu, L L^T) [U: normal_fullrank].
    Same Philox draws on both sides: the iterates (mu, packed L), the ELBO estimate, the amplitude summaries and the
    posterior matrix must agree with the CPU restatement."""
    p = make_problem(120, 14, 5, lik, seed=21)
    o = orc.Oracle(p["phi"], p["y"], lik, nthreads=2)
    D = o.D
    nL = D * (D + 1) // 2
    rng = np.random.default_rng(4)
    mu0 = rng.normal(0, 0.1, D)
    f = _fit(p["phi"], p["y"], lik, prec, eng, algorithm="fullrank", seed=77, n_draws=64)
    f.set_variational(mu0)                                     # L = I
    assert np.array_equal(f.get_cholesky(), orc.pack_tril(np.eye(D)))
    mu, L, hist = mu0.copy(), orc.pack_tril(np.eye(D)), np.zeros(D + nL)
    tol = 1e-9 if prec == "fp64" else 2e-4
    f.advi_steps(7, 0.1, reset=True)
    for it in range(1, 8):
        assert o.fr_step(mu, L, hist, 0.1, it, 77, it - 1) == 0
    gm, _ = f.get_variational()
    assert rel_err(gm, mu) <= tol and rel_err(f.get_cholesky(), L) <= tol
    f.advi_steps(5, 0.1)                                        # continues: it = 8.., draws 7..
    for it in range(8, 13):
        o.fr_step(mu, L, hist, 0.1, it, 77, it - 1)
    gm, gom = f.get_variational()
    gL = f.get_cholesky()
    assert rel_err(gm, mu) <= 3 * tol and rel_err(gL, L) <= 3 * tol
    diag = orc.unpack_tril(gL, D).diagonal()
    assert np.allclose(gom, np.log(np.abs(diag)), rtol=1e-12)     # get_variational reports omega = log |L_ii|
    # ELBO / summaries / posterior at exactly the oracle's state
    f.set_variational(mu)
    f.set_cholesky(L)
    e = f.elbo(23)[0]     # (the state after 12 steps from L = I is far out: log densities ~ -1e12, hence 2e-5 for the bf16-split engine)
    assert e == pytest.approx(o.fr_elbo(mu, L, 23, 77, 0), rel=1e-10 if prec == "fp64" else 2e-5)
    assert f.elbo(8)[0] == pytest.approx(o.fr_elbo(mu, L, 8, 77, 23), rel=1e-10 if prec == "fp64" else 2e-5)   # counter advanced
    got, ref = f.summarize(), o.summary_fullrank(mu, L, 64, 77)
    assert np.allclose(got[:, :7], ref[:, :7], rtol=1e-9) and np.array_equal(got[:, 7], ref[:, 7])
    P = f.posterior(11)
    cols = f.posterior_columns()
    ia = cols.index("amplitude[1]")
    assert P.shape == (11, len(cols)) and np.allclose(P[:, ia:ia + 14], f.draws()[:11], rtol=1e-12)
    Lf = orc.unpack_tril(L, D)
    eta = np.array([[orc.normal(77, 3, d, i) for i in range(D)] for d in range(11)])
    zeta = mu[None] + eta @ Lf.T
    assert np.allclose(P[:, 2], -0.5 * (eta ** 2).sum(1), rtol=1e-12)
    lp_ref = np.array([o.log_prob(z, True, False, grad=False) for z in zeta])
    assert np.allclose(P[:, 1], lp_ref, rtol=1e-10 if prec == "fp64" else 5e-5)
    f.close()


def test_fullrank_full_fit_and_errors():
    """whole path with algorithm = "fullrank": L-BFGS init -> adapt -> SGA -> draws (fp64 engine reproduces the oracle's
    decisions), plus the refusals: gene sharding, pathfinder"""
    p = make_problem(150, 12, 4, "gaussian", seed=9)
    o = orc.Oracle(p["phi"], p["y"], "gaussian", nthreads=2)
    cfg = dict(n_iter=300, mle_iter=50, elbo_samples=20, n_draws=200, seed=312)
    f = bvg.Fit(p["phi"], p["y"], "gaussian", "fp64", "simt", algorithm="fullrank", **cfg)
    got = f.run()
    st = f.stats()
    th, _ = o.lbfgs(max_iter=50)
    mu, L, info = o.advi_fullrank(th, iter=300, elbo_samples=20, seed=312)
    assert info["rc"] == 0 and st["eta"] == info["eta"] and st["advi_iters"] == info["iters"] and st["converged"] == info["converged"]
    assert st["elbo"] == pytest.approx(info["elbo"], rel=1e-6)
    ref = o.summary_fullrank(mu, L, 200, 312)
    assert np.allclose(got[:, :7], ref[:, :7], rtol=1e-5)
    f.close()
    with pytest.raises(bvg._lib.BvgError, match="gene-sharded"):
        bvg.Fit(p["phi"], p["y"], "gaussian", "fast", "auto", algorithm="fullrank", world_size=2, rank=0, G_total=24, gene_offset=0)
    with pytest.raises(ValueError, match="valid variational inference"):
        bvg.Fit(p["phi"], p["y"], "gaussian", "fast", "auto", algorithm="hmc")
    fm = bvg.Fit(p["phi"], p["y"], "gaussian", "fast", "auto")
    with pytest.raises(bvg._lib.BvgError, match="fullrank"):
        fm.get_cholesky()
    fm.close()


@pytest.mark.parametrize("lik", ["gaussian", "nb"])
@pytest.mark.parametrize("prec,eng", ENGINES)
def test_pathfinder_single_path_matches_oracle(lik, prec, eng):
    """algorithm = "pathfinder" (R/findSpatiallyVariableFeaturesBayes.R:366-375) [U: pathfinder_lbfgs_single]: over the first
    L-BFGS iterates (before the two optimisers' round-off separates the paths) the per-iterate ELBO estimates, the chosen
    iterate, its draws and their log p - log q must equal the CPU restatement's (same Philox draws)."""
    p = make_problem(150, 12, 4, lik, seed=9)
    o = orc.Oracle(p["phi"], p["y"], lik, nthreads=2)
    kw = dict(max_lbfgs_iters=9, history=5, num_elbo_draws=12, num_draws=40, seed=21)
    ramp, rlr, rinfo = o.pf_single(1, **kw)
    f = _fit(p["phi"], p["y"], lik, prec, eng, algorithm="pathfinder", seed=21, elbo_samples=12, lbfgs_history=5,
             pf_max_lbfgs_iters=9, pf_single_draws=40)
    amp, lr, info = f.pathfinder_single(1)
    assert rinfo["rc"] == 0 and info["lbfgs_iters"] == rinfo["lbfgs_iters"] == 9
    tol = 1e-7 if prec == "fp64" else 2e-3
    e, re = info["elbos"][1:10], rinfo["elbos"][1:10]
    assert np.all(np.isfinite(e) == np.isfinite(re))
    ok = np.isfinite(re)
    assert np.allclose(e[ok], re[ok], rtol=tol)
    if prec == "fp64":
        assert info["best_iter"] == rinfo["best_iter"] and info["n_pairs"] == rinfo["n_pairs"]
        assert np.allclose(amp, ramp, rtol=1e-6) and np.allclose(lr, rlr, rtol=1e-7, atol=1e-6)
    f.close()


def test_pathfinder_full_run_psis_and_summary():
    """multi-path Pathfinder through bvg_fit_run: PSIS resampling + amplitude summaries; with few L-BFGS iterations the
    fp64 engine reproduces the oracle's run exactly (paths, resampled indices, summary table)"""
    p = make_problem(150, 12, 4, "gaussian", seed=9)
    o = orc.Oracle(p["phi"], p["y"], "gaussian", nthreads=2)
    kw = dict(max_lbfgs_iters=8, history=5, num_elbo_draws=10, num_draws=60, num_paths=3, num_psis_draws=90, seed=5)
    ref, rinfo = o.pathfinder(**kw)
    assert rinfo["rc"] == 0
    f = bvg.Fit(p["phi"], p["y"], "gaussian", "fp64", "simt", algorithm="pathfinder", seed=5, elbo_samples=10, lbfgs_history=5,
                pf_max_lbfgs_iters=8, pf_single_draws=60, pf_num_paths=3, n_draws=90)
    got = f.run()
    tr, st = f.trace(), f.stats()
    assert len(tr) == 3 and [int(r[1]) for r in tr] == [q["best_iter"] for q in rinfo["paths"]]
    assert np.allclose(tr[:, 2], [q["best_elbo"] for q in rinfo["paths"]], rtol=1e-7)
    assert st["eta"] == pytest.approx(rinfo["khat"], rel=1e-5)          # Pareto k of the importance ratios
    assert np.allclose(got[:, :7], ref[:, :7], rtol=1e-6) and np.array_equal(got[:, 7], ref[:, 7])
    assert f.draws().shape == (90, 12)
    f.close()
    # a longer run on the fast engine: finite, positive, and the selected iterate is an interior one
    f = bvg.Fit(p["phi"], p["y"], "gaussian", "fast", "auto", algorithm="pathfinder", seed=5, pf_max_lbfgs_iters=60, elbo_samples=20,
                pf_single_draws=100, n_draws=150)
    got = f.run()
    tr = f.trace()
    assert np.all(np.isfinite(got[:, :7])) and np.all(got[:, 0] > 0) and len(tr) == 2 and np.all(tr[:, 1] >= 1)
    f.close()
    with pytest.raises(bvg._lib.BvgError, match="gene-sharded"):
        bvg.Fit(p["phi"], p["y"], "gaussian", "fast", "auto", algorithm="pathfinder", world_size=2, rank=0, G_total=24, gene_offset=0)


def test_divergent_fit_fails_like_the_reference_algorithm():
    """a problem on which eta adaptation picks 1.0 and stochastic gradient ascent then overflows:
    Stan raises, the oracle returns -4, the engine must return BVG_ERR_NUMERIC (not garbage)"""
    p = make_problem(400, 120, 8, "gaussian", seed=31)
    o = orc.Oracle(p["phi"], p["y"], "gaussian", nthreads=4)
    th, _ = o.lbfgs(max_iter=100)
    _, _, info = o.advi(th, iter=400, elbo_samples=30, seed=312)
    assert info["rc"] == -4 and info["eta"] == 1.0
    f = bvg.Fit(p["phi"], p["y"], "gaussian", "fp64", "simt", n_iter=400, mle_iter=100, elbo_samples=30, seed=312)
    with pytest.raises(bvg._lib.BvgError) as e:
        f.run()
    assert e.value.code == -4 and f.stats()["eta"] == 1.0
    f.close()


def test_error_paths():
    p = make_problem(40, 5, 3)
    with pytest.raises(ValueError):
        bvg.Fit(p["phi"], p["y"], "poisson")
    with pytest.raises(bvg._lib.BvgError) as e:
        bvg.Fit(p["phi"], p["y"], "gaussian", "fp64", "tcgen05")
    assert "fp64" in str(e.value)
    f = bvg.Fit(p["phi"], p["y"])
    with pytest.raises(bvg._lib.BvgError):
        f.summary()
    with pytest.raises(ValueError):
        f.log_prob(np.zeros(3))
    with pytest.raises(bvg._lib.BvgError):
        bvg.Fit(np.linalg.qr(np.random.default_rng(0).normal(size=(600, 511)))[0], np.zeros((600, 2)))  # k > 510
    with pytest.raises(bvg._lib.BvgError):
        bvg.Fit(p["phi"], -np.ones((40, 5), dtype=np.int32), "nb")


def test_full_size_properties_cfg2():
    """BASELINE config 2 shape (4,992 spots x 3,000 genes, k = 20): size-independent checks.
    (1) orthonormal-Phi identity: SSE from sufficient statistics equals the dense pass;
    (2) directional derivative: central difference of lp along a random direction equals grad.v."""
    from bayesvg_b200 import synth
    ds = synth.dataset("cfg2", lambda X, C, ls, kern, nu: bvg.basis_build(X, C, ls, kern, nu))
    phi, Y = ds["phi"], ds["Y"]
    M, G, k = 4992, 3000, 20
    lay = orc.layout("gaussian", G, k)
    rng = np.random.default_rng(0)
    th = rng.normal(0, 0.2, lay["D"])
    g64 = None
    for prec, eng, tol in (("fp64", "simt", 1e-9), ("fast", "auto", 2e-5)):
        f = bvg.Fit(phi, Y, "gaussian", prec, eng)
        lp, g = f.log_prob(th, True, True)
        # closed form through S2, S1, c = Phi'y, o = Phi'1
        beta = th[0] + np.exp(th[1]) * th[2:2 + G]
        A = np.exp(2 * th[lay["amplitude"]:lay["amplitude"] + G])
        z = th[lay["z_alpha_t"]:lay["z_alpha_t"] + k * G].reshape(G, k).T
        alpha = th[lay["mu_alpha"]:lay["mu_alpha"] + k][:, None] + np.exp(th[lay["sigma_alpha"]:lay["sigma_alpha"] + k])[:, None] * z
        S2, S1, c, o1 = (Y ** 2).sum(0), Y.sum(0), phi.T @ Y, phi.sum(0)
        sse = (S2 - 2 * beta * S1 - 2 * A * (c * alpha).sum(0) + M * beta ** 2 + 2 * beta * A * (o1 @ alpha)
               + A ** 2 * (alpha ** 2).sum(0)).sum()
        sy = np.exp(th[lay["sigma_y"]])
        u = th[lay["amplitude"]:lay["amplitude"] + G]
        ma, lsa = th[lay["mu_amplitude"]], th[lay["sigma_amplitude"]]
        mua, lsal = th[lay["mu_alpha"]:lay["mu_alpha"] + k], th[lay["sigma_alpha"]:lay["sigma_alpha"] + k]
        ref = (-M * G * th[lay["sigma_y"]] - 0.5 * sse / sy ** 2 - 0.5 * (z ** 2).sum() - 0.5 * (th[2:2 + G] ** 2).sum()
               - th[0] ** 2 / 8 - 0.5 * np.exp(2 * th[1]) - ma ** 2 / 8 - 0.5 * np.exp(2 * lsa) - G * lsa - u.sum()
               - 0.5 * ((u - ma) ** 2).sum() / np.exp(2 * lsa) - (mua ** 2).sum() / 8 - 0.5 * np.exp(2 * lsal).sum()
               - 0.5 * sy ** 2 + th[1] + lsa + u.sum() + lsal.sum() + th[lay["sigma_y"]])
        assert abs(lp - ref) <= tol * abs(ref), (prec, lp, ref)
        if prec == "fp64":
            g64 = g
            v = rng.normal(size=lay["D"])
            v /= np.linalg.norm(v)
            h = 1e-4
            fd = (f.log_prob(th + h * v, grad=False) - f.log_prob(th - h * v, grad=False)) / (2 * h)
            assert fd == pytest.approx(float(g @ v), rel=1e-5)
        else:
            assert rel_err(g, g64) <= 2e-4
        f.close()


@pytest.mark.parametrize("name,M,G,k,lik,kern,nu", [
    ("cfg2", 4992, 3000, 20, "gaussian", "matern", 2.5),      # BASELINE.json configs[1]: the shape the metric is quoted on
    ("cfg3", 4992, 3000, 20, "nb", "exp_quad", 0.0),          # configs[2]: counts, negative-binomial, exp_quad
    ("cfg4_shard", 50000, 1250, 20, "gaussian", "matern", 1.5),  # configs[3]: what each of 8 GPUs holds at 50k x 10k
    ("k400", 20000, 256, 400, "gaussian", "matern", 1.5),     # configs[4]'s basis size on the large-basis tcgen05 kernels
])
def test_full_size_log_prob_and_gradient_against_the_oracle(name, M, G, k, lik, kern, nu):
    """The engines the bench quotes, at the shapes it quotes them on, against the CPU ORACLE (not against each other):
    log density and full gradient at a random unconstrained vector.  north_star gates: <= 1e-6 (fp64), <= 1e-3 (fast path);
    asserted tighter."""
    from bayesvg_b200 import synth
    from bayesvg_b200.api import length_scale_kmeans
    X = synth.scale_coords(synth.visium_grid() if M == 4992 else synth.random_points(M, 1))
    C = synth.det_kmeans(X, k, 312, iters=30) if k <= 30 else bvg.kmeans_centers(X, k, 30, 1, rng=312)
    phi = bvg.basis_build(X, C, length_scale_kmeans(C), kern, nu)
    Y, _ = synth.expression(phi, G, lik, 312)
    if lik == "nb":
        Y = np.asfortranarray(Y.astype(np.int32))
    o = orc.Oracle(phi, np.asfortranarray(Y, dtype=np.float64), lik, nthreads=os.cpu_count() or 4)
    th = np.random.default_rng(5).normal(0, 0.2, o.D)
    rlp, rg = o.log_prob(th, True, True)
    engines = [("fast", "tcgen05", 2e-5, 2e-4)]
    if name in ("cfg2", "cfg3"):
        engines.append(("fp64", "simt", 1e-10, 1e-9))
    for prec, eng, tol_lp, tol_g in engines:
        f = bvg.Fit(phi, Y, lik, prec, eng)
        assert f.stats()["engine_used"] == (2 if eng == "tcgen05" else 1)
        lp, g = f.log_prob(th, True, True)
        assert abs(lp - rlp) <= tol_lp * abs(rlp), (name, prec, lp, rlp)
        assert rel_err(g, rg) <= tol_g, (name, prec, rel_err(g, rg))
        if prec == "fast" and k <= 23:
            # ... and ten ADVI iterations through the persistent kernel against the oracle's fp64 loop (same Philox draws)
            mu, om, hist = np.zeros(o.D), np.zeros(o.D), np.zeros(2 * o.D)
            f.set_variational(mu, None)
            f.advi_steps(0, 0.1, reset=True)
            for it in range(1, 11):
                o.advi_step(mu, om, hist, 0.1, it, 312, it - 1)
            f.advi_steps(10, 0.1)
            gmu, gom = f.get_variational()
            assert rel_err(gmu, mu) < 1e-3 and rel_err(gom, om) < 1e-3, (name, rel_err(gmu, mu), rel_err(gom, om))
        f.close()


def test_fast_lbfgs_hands_over_to_the_fp64_evaluator_at_its_noise_floor():
    """BASELINE configs[2] (nb, exp_quad, 4,992 x 3,000): the fast path's line search fails after ~350 of the 1,000 L-BFGS
    iterations (its evaluator's noise, scripts/lsfail_probe.py); lbfgs_run then continues from that point with the fp64
    evaluator (fp64_shadow.cu) and must end like the all-fp64 run: same termination after the same 1,000 iterations, a log
    density no worse than the all-fp64 run's (both stop at the iteration cap, on slightly different paths), and the fast engine
    back in place afterwards."""
    from bayesvg_b200 import synth
    ds = synth.dataset("cfg3", lambda X, C, ls, kern, nu: bvg.basis_build(X, C, ls, kern, nu))
    f = bvg.Fit(ds["phi"], ds["Y"], "nb", "fast", "tcgen05", seed=312)
    th = f.mle()
    st = f.stats()
    fd = bvg.Fit(ds["phi"], ds["Y"], "nb", "fp64", "simt", seed=312)
    fd.mle()
    sd = fd.stats()
    assert st["mle_fp64_from"] > 0 and st["mle_term"] == sd["mle_term"] and st["mle_iters"] == sd["mle_iters"] == 1000
    assert st["mle_lp"] >= sd["mle_lp"] - 2e-5 * abs(sd["mle_lp"])   # round 1 (no hand-over): 1,489 = 9.4e-5 below
    assert st["engine_used"] == 2
    lp_fast = f.log_prob(th, False, False, grad=False)          # evaluated by the tcgen05 engine again
    lp_64 = fd.log_prob(th, False, False, grad=False)
    assert lp_fast == pytest.approx(lp_64, rel=1e-6)
    assert f.advi_steps(10, 0.1, reset=True) > 0                 # ... and the persistent kernel runs on the restored buffers
    f.close()
    fd.close()


@pytest.mark.parametrize("lik,eng,M,G,k", [("gaussian", "tcgen05", 700, 300, 20), ("nb", "tcgen05", 700, 300, 12),
                                             ("gaussian", "auto", 640, 45, 64), ("nb", "simt", 300, 70, 7)])
def test_fp64_hand_over_of_the_fast_lbfgs_on_every_fast_engine(lik, eng, M, G, k):
    """fp64_shadow.cu behind every fast engine (fused tcgen05, large-basis tcgen05, SIMT fp32), forced at iteration 6
    (BVG_MLE_FP64_AT): the run finishes on fp64 evaluations (Y widened from the resident fp32 copy), reaches the all-fp64 run's
    optimum, and the fast engine evaluates again afterwards."""
    p = make_problem(M, G, k, lik, seed=41)
    os.environ["BVG_MLE_FP64_AT"] = "6"
    try:
        f = bvg.Fit(p["phi"], p["y"], lik, "fast", eng, seed=3, mle_iter=60)
        th = f.mle()
        st = f.stats()
    finally:
        del os.environ["BVG_MLE_FP64_AT"]
    fd = bvg.Fit(p["phi"], p["y"], lik, "fp64", "simt", seed=3, mle_iter=60)
    fd.mle()
    sd = fd.stats()
    assert st["mle_fp64_from"] == 6 and st["mle_iters"] == 60 and sd["mle_fp64_from"] == 0
    # (the forced hand-over clears the curvature history at iteration 6, so after 60 iterations the two runs are at different
    # stages of the same climb: the run must have climbed, not matched)
    lp0 = fd.log_prob(np.zeros(fd.D), False, False, grad=False)
    assert st["mle_lp"] > lp0 + 0.5 * (sd["mle_lp"] - lp0)
    # the reported optimum is an fp64 value of the point returned, and the fast engine is back for everything that follows
    assert fd.log_prob(th, False, False, grad=False) == pytest.approx(st["mle_lp"], rel=1e-6)
    assert f.log_prob(th, False, False, grad=False) == pytest.approx(st["mle_lp"], rel=2e-4)
    assert f.stats()["engine_used"] == (1 if eng == "simt" else 2)
    f.advi()
    assert np.all(np.isfinite(f.summarize()))
    f.close()
    fd.close()


def test_cfg1_fit_agrees_with_the_oracle_fit_fixture(golden_dir):
    """north_star's second check at the scale it names: the posterior amplitude summaries of the whole default fit (L-BFGS
    init -> eta adaptation -> 3,000 SGA iterations -> 1,000 draws) on the stand-in for configs[0] (2,444 spots x 3,000 HVGs,
    Matern-2.5, gaussian) against the CPU oracle's fit of the same data with the same Philox streams, stored in
    tests/golden/cfg1_fit.npz (tests/golden/make_cfg1_fit_fixture.py).  Stated Monte Carlo tolerance: Spearman >= 0.99 and
    top-1,000 overlap >= 0.98 for the fast path (>= 0.999 / 0.99 in fp64)."""
    from scipy.stats import spearmanr

    from bayesvg_b200 import synth
    fx = np.load(os.path.join(golden_dir, "cfg1_fit.npz"))
    ds = synth.dataset("cfg1", lambda X, C, ls, kern, nu: bvg.basis_build(X, C, ls, kern, nu))
    assert float(np.abs(ds["Y"]).sum()) == pytest.approx(float(fx["y_checksum"]), rel=1e-9)     # same data as the fixture's
    assert float(np.abs(ds["phi"]).sum()) == pytest.approx(float(fx["phi_checksum"]), rel=1e-8)
    ref = fx["amplitude_mean"]
    for prec, eng, rho_min, top_min in (("fast", "tcgen05", 0.99, 0.98), ("fp64", "simt", 0.999, 0.99)):
        f = bvg.Fit(ds["phi"], ds["Y"], "gaussian", prec, eng, seed=312)
        tab = f.run()
        st = f.stats()
        f.close()
        assert st["eta"] == float(fx["eta"]) and st["advi_iters"] == int(fx["advi_iters"])
        rho = float(spearmanr(tab[:, 0], ref).correlation)
        top = len(set(np.argsort(-tab[:, 0])[:1000]) & set(np.argsort(-ref)[:1000])) / 1000
        assert rho >= rho_min and top >= top_min, (prec, rho, top)
        if prec == "fast":
            assert st["mle_iters"] == 1000 and st["mle_term"] == 5   # the fast path's L-BFGS runs its 1,000 iterations like the fp64 one


def _gmm_data(N=150, D=4, K=3, seed=0):
    rng = np.random.default_rng(seed)
    cent = rng.normal(0, 3, (K, D))
    lab = rng.integers(0, K, N)
    return cent[lab] + rng.normal(0, 0.7, (N, D)), lab


def test_gmm_log_prob_matches_oracle():
    """inst/clusterSVGs.stan (the program behind clusterSVGsBayes, R/clusterSVGsBayes.R:139-160): log density and gradient at
    unconstrained vectors, all (jacobian, propto) combinations, against the CPU restatement"""
    for (N, D, K) in ((150, 4, 3), (97, 30, 5), (40, 2, 2), (500, 7, 16)):
        X, _ = _gmm_data(N, D, K, seed=N)
        o = orc.GmmOracle(X, K)
        rng = np.random.default_rng(1)
        for scale in (0.0, 0.3, 1.0):
            th = rng.normal(0, 1, o.D) * scale
            for jac in (0, 1):
                for pro in (0, 1):
                    lp, g = bvg.engine.gmm_log_prob(X, K, th, jac, pro)
                    rlp, rg = o.log_prob(th, jac, pro)
                    assert lp == pytest.approx(rlp, rel=1e-11) and rel_err(g, rg) <= 1e-11, (N, D, K, scale, jac, pro)
                    assert bvg.engine.gmm_log_prob(X, K, th, jac, pro, grad=False) == pytest.approx(rlp, rel=1e-11)


@pytest.mark.parametrize("name", ["gmm_small", "gmm_default_shape"])
def test_gmm_log_prob_golden(golden_dir, name):
    """the device evaluation of inst/clusterSVGs.stan against the committed torch-autograd transcription"""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    X, K = z["X"], int(z["K"])
    for jac in (0, 1):
        for pro in (0, 1):
            for t, th in enumerate(z["thetas"]):
                lp, g = bvg.engine.gmm_log_prob(X, K, th, jac, pro)
                assert lp == pytest.approx(z[f"lp_j{jac}_p{pro}"][t], rel=1e-11), (jac, pro, t)
                assert rel_err(g, z[f"grad_j{jac}_p{pro}"][t]) <= 1e-10, (jac, pro, t)


@pytest.mark.parametrize("algorithm", ["meanfield", "fullrank"])
def test_gmm_fit_matches_oracle(algorithm):
    """the whole clusterSVGsBayes() fit (eta adaptation, SGA, ELBO test, draws, soft assignments, log-likelihood) on the device
    against the oracle driving the same algorithm with the same Philox streams"""
    X, _ = _gmm_data(150, 4, 3, seed=2)
    o = orc.GmmOracle(X, 3)
    kw = dict(iter=400, elbo_samples=40, seed=17)
    fr = algorithm == "fullrank"
    mu, var, info = (o.advi_fullrank if fr else o.advi)(np.zeros(o.D), **kw)
    assert info["rc"] == 0
    res = bvg.engine.gmm_fit(X, 3, algorithm, n_iter=400, n_draws=64, elbo_samples=40, seed=17)
    i = res["info"]
    assert (i["eta"], i["iters"], i["converged"]) == (info["eta"], info["iters"], info["converged"])
    assert i["elbo"] == pytest.approx(info["elbo"], rel=1e-7)
    assert rel_err(res["mu"], mu) <= 1e-7 and rel_err(res["var"], var) <= 1e-7
    draws, soft, ll = o.post(mu, var, fr, 64, 17)
    assert np.allclose(res["draws"], draws, rtol=1e-6, atol=1e-9) and np.allclose(res["soft"], soft, atol=1e-7)
    assert i["log_likelihood"] == pytest.approx(ll, rel=1e-7)
    assert np.allclose(res["soft"].sum(1), 1.0) and np.allclose(res["draws"][:, :3].sum(1), 1.0)     # simplex draws


def test_gmm_batch_evaluation_and_pathfinder_match_the_oracle():
    """clusterSVGsBayes(algorithm = "pathfinder") (R/clusterSVGsBayes.R:162-170): bvg_gmm_eval -- B points per launch, one CTA
    each -- against the oracle's log density / gradient, then a whole Pathfinder path with the DEVICE evaluator against the
    oracle's C restatement of Pathfinder on the same program (same Philox streams), and the posterior pass for given draws."""
    from bayesvg_b200 import gmm_pathfinder as gpf
    X, lab = _gmm_data(150, 4, 3, seed=2)
    o = orc.GmmOracle(X, 3)
    ev = gpf.DeviceEvaluator(X, 3)
    rng = np.random.default_rng(0)
    ths = rng.normal(0, 0.4, (37, o.D))
    lps = ev.lp_batch(ths)
    ref = np.array([o.log_prob(t, True, True, grad=False) for t in ths])
    assert np.allclose(lps, ref, rtol=1e-11)
    lp, g = ev.lp_grad(ths[5])
    rlp, rg = o.log_prob(ths[5], True, True)
    assert lp == pytest.approx(rlp, rel=1e-11) and rel_err(g, rg) < 1e-10
    kw = dict(max_lbfgs_iters=30, history=6, num_elbo_draws=8, num_draws=40, seed=11)
    phi_o, lr_o, io = o.pf_single_phi(1, **kw)
    phi_p, lr_p, ip = gpf.single_path(ev, 1, 11, 40, 30, 6, 8, 0.01)
    assert io["rc"] == 0 and (io["lbfgs_iters"], io["best_iter"], io["n_pairs"]) == (ip["lbfgs_iters"], ip["best_iter"], ip["n_pairs"])
    assert ip["best_elbo"] == pytest.approx(io["best_elbo"], rel=1e-8)
    assert rel_err(phi_p, phi_o) < 1e-7 and np.allclose(lr_p, lr_o, atol=1e-5)
    # posterior pass for given unconstrained draws: a full-rank family with zero scale draws its mean -> compare with k_gmm_post's twin
    draws, soft, ll = ev.posterior(phi_p)
    assert draws.shape == (40, 3 + 2 * 3 * 4) and np.allclose(draws[:, :3].sum(1), 1.0) and np.allclose(soft.sum(1), 1.0)
    d1, s1, ll1 = o.post(phi_p[0], np.full(o.D, -800.0), False, 3, 17)      # exp(-800) = 0: every draw is the point itself
    dd, ss, l2 = ev.posterior(phi_p[:1])
    assert np.allclose(dd[0], d1[0], rtol=1e-9) and np.allclose(ss, s1, atol=1e-9) and l2 == pytest.approx(ll1, rel=1e-9)
    ev.close()
    # the whole thing through the host mirror: planted modules recovered
    res = gpf.gmm_pathfinder(X, 3, n_draws=100, elbo_samples=20, seed=5, num_paths=2, max_lbfgs_iters=60, history=20)
    assert res["draws"].shape == (100, 27) and np.isfinite(res["info"]["log_likelihood"]) and np.allclose(res["soft"].sum(1), 1.0)
    hard = res["soft"].argmax(1)
    purity = sum(np.bincount(hard[lab == j], minlength=3).max() for j in range(3)) / len(lab)
    assert purity >= 0.9, purity


def test_cluster_svgs_bayes_host_mirror():
    """clusterSVGsBayes() (R/clusterSVGsBayes.R:63-319; test_bayesVG.R:114-118, :188-192): list with cluster_df, pca_embedding,
    model_fit, log_likelihood, BIC; planted gene modules are recovered up to a relabelling"""
    rng = np.random.default_rng(5)
    M, n_per, K = 400, 40, 3
    patterns = rng.normal(size=(K, M))
    lab = np.repeat(np.arange(K), n_per)
    data = patterns[lab] * rng.uniform(0.8, 1.5, (K * n_per, 1)) + 0.4 * rng.normal(size=(K * n_per, M))
    genes = [f"g{i}" for i in range(K * n_per)]
    obj = bvg.SpatialData(rng.uniform(0, 10, (M, 2)), data=data, genes=genes)
    res = bvg.clusterSVGsBayes(obj, svgs=genes, n_clust=K, n_PCs=5, n_iter=3000, n_draws=200, verbose=False)
    assert set(res) == {"pca_embedding", "cluster_df", "model_fit", "log_likelihood", "BIC"}
    cdf = res["cluster_df"]
    assert list(cdf.columns) == ["gene"] + [f"prob_cluster_{j}" for j in (1, 2, 3)] + ["assigned_cluster"] and len(cdf) == K * n_per
    assert res["pca_embedding"].shape == (K * n_per, 5) and isinstance(res["model_fit"], bvg.VariationalFit)
    assert np.isfinite(res["log_likelihood"]) and res["BIC"] == pytest.approx(-2 * res["log_likelihood"] + (K - 1 + 2 * K * 5) * np.log(K * n_per))
    got = cdf["assigned_cluster"].to_numpy()
    purity = sum(np.bincount(got[lab == j], minlength=K + 1).max() for j in range(K)) / len(lab)
    assert purity >= 0.9, purity
    assert list(res["model_fit"].draws("theta").columns) == ["theta[1]", "theta[2]", "theta[3]"]
    rp = bvg.clusterSVGsBayes(obj, svgs=genes, n_clust=K, n_PCs=5, n_draws=200, algorithm="pathfinder", verbose=False)
    gotp = rp["cluster_df"]["assigned_cluster"].to_numpy()
    # (Pathfinder's L-BFGS paths start next to the symmetric point init = 0.01 and may end in a local mode of the mixture
    # posterior that splits one module -- 0.88 on this data; the check is that the driver runs and clusters, not that it is ADVI)
    assert sum(np.bincount(gotp[lab == j], minlength=K + 1).max() for j in range(K)) / len(lab) >= 0.6
    assert np.isfinite(rp["BIC"]) and rp["model_fit"].draws("theta").shape == (200, 3)
    with pytest.raises(ValueError, match="valid number of clusters"):
        bvg.clusterSVGsBayes(obj, svgs=genes, n_clust=1)
    with pytest.raises(ValueError, match="must be supplied"):
        bvg.clusterSVGsBayes(obj)

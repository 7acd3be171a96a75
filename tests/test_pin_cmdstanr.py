"""The cmdstanr pin kit (tests/golden/pin, scripts/pin_with_cmdstanr.R).

CPU part: the kit's inputs say exactly what the golden fixtures say (so whoever runs the R script evaluates the reference at the
points the rest of the suite uses), and -- once `<name>.cmdstanr.json` files are present -- the ORACLE equals cmdstanr's
`$log_prob()` / `$grad_log_prob()` to 1e-8 relative (north_star: <= 1e-6 in fp64).  GPU part: the CUDA engine against the same
files.  Without the R outputs the comparisons skip and parity stays "unpinned" (DESIGN.md section 6)."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIN = os.path.join(ROOT, "tests", "golden", "pin")
NAMES = [e["name"] for e in json.load(open(os.path.join(PIN, "index.json")))]
LIK = {"gp": "gaussian", "gp2": "gaussian", "gp3": "nb", "gp4": "nb"}


def _kit(name):
    data = json.load(open(os.path.join(PIN, f"{name}.data.json")))
    th = json.load(open(os.path.join(PIN, f"{name}.theta.json")))
    M, G = data["M"], data["G"]
    y = np.asfortranarray(np.array(data["y"], dtype=np.float64).reshape(G, M).T)
    return data, np.array(data["phi"]), y, np.array(th["thetas"]), th["program"]


@pytest.mark.parametrize("name", NAMES)
def test_kit_inputs_equal_the_golden_fixtures(name):
    data, phi, y, thetas, program = _kit(name)
    d = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
    assert np.array_equal(phi, d["phi"]) and np.array_equal(y, d["y"]) and np.array_equal(thetas, d["thetas"])
    M, G = data["M"], data["G"]
    # long format of R/findSpatiallyVariableFeaturesBayes.R:293-331: gene-major, 1-based ids, N = M G
    assert data["N"] == M * G and data["spot_id"] == list(range(1, M + 1)) * G
    assert data["gene_id"] == [g + 1 for g in range(G) for _ in range(M)]
    assert program == {"gp": "approxGP.stan", "gp2": "approxGP2.stan", "gp3": "approxGP3.stan", "gp4": "approxGP4.stan"}[name.split("_")[0]]
    if "gene_depths" in d.files:
        assert np.array_equal(np.array(data["gene_depths"]), d["gene_depths"])
    # ... and the fixture's own numbers are what the oracle gives at these inputs
    o = orc.Oracle(phi, y, LIK[name.split("_")[0]], gene_depths=(np.array(data["gene_depths"]) if "gene_depths" in data else None))
    for t, th in enumerate(thetas):
        lp, g = o.log_prob(th, True, True)
        assert lp == pytest.approx(float(d["lp_j1_p1"][t]), rel=1e-10)
        assert np.linalg.norm(g - d["grad_j1_p1"][t]) <= 1e-9 * np.linalg.norm(g)


def _cmdstanr(name):
    fn = os.path.join(PIN, f"{name}.cmdstanr.json")
    if not os.path.exists(fn):
        pytest.skip("no cmdstanr output for this fixture yet: run scripts/pin_with_cmdstanr.R (parity stays unpinned)")
    return json.load(open(fn))


def _compare(evaluate, name):
    ref = _cmdstanr(name)
    data, phi, y, thetas, _ = _kit(name)
    for t, th in enumerate(thetas):
        for jac in (0, 1):
            g_ref = np.array(ref[f"grad_j{jac}"][t], dtype=np.float64)
            lp_ref = float(np.atleast_1d(ref[f"lp_j{jac}"])[t])
            # cmdstanr's log_prob drops or keeps the constants depending on its version: the gradient is the same either way
            lp1, g = evaluate(th, bool(jac), True)
            lp0, _ = evaluate(th, bool(jac), False)
            assert np.linalg.norm(g - g_ref) <= 1e-8 * np.linalg.norm(g_ref), (name, t, jac)
            assert min(abs(lp1 - lp_ref), abs(lp0 - lp_ref)) <= 1e-8 * abs(lp_ref), (name, t, jac, lp1, lp0, lp_ref)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_equals_cmdstanr(name):
    data, phi, y, _, _ = _kit(name)
    o = orc.Oracle(phi, y, LIK[name.split("_")[0]], gene_depths=(np.array(data["gene_depths"]) if "gene_depths" in data else None))
    _compare(lambda th, jac, propto: o.log_prob(th, jac, propto), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_engine_equals_cmdstanr(name):
    import bayesvg_b200 as bvg
    data, phi, y, _, _ = _kit(name)
    f = bvg.Fit(phi, y, LIK[name.split("_")[0]], "fp64", "simt", gene_depths=(np.array(data["gene_depths"]) if "gene_depths" in data else None))
    try:
        _compare(lambda th, jac, propto: f.log_prob(th, jac, propto), name)
    finally:
        f.close()


def test_kit_is_complete():
    have = {os.path.basename(p) for p in glob.glob(os.path.join(PIN, "*.json"))}
    for n in NAMES:
        assert f"{n}.data.json" in have and f"{n}.theta.json" in have
    assert os.path.exists(os.path.join(ROOT, "scripts", "pin_with_cmdstanr.R"))

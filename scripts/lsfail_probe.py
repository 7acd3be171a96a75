"""python scripts/lsfail_probe.py [cfg3|cfg1] -- where the fast path's L-BFGS stops, what does the evaluated log density look like
along the steepest-descent direction?  Prints f(x + a g) - f(x) against the first-order prediction a |g|^2 for the fast
(tcgen05) and the fp64 engine at the point the fast L-BFGS returned."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from bayesvg_b200 import synth  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    ds = synth.dataset(name, lambda X, C, ls, kern, nu: bvg.basis_build(X, C, ls, kern, nu))
    phi, Y, lik = ds["phi"], ds["Y"], ds["likelihood"]
    ff = bvg.Fit(phi, Y, lik, "fast", "tcgen05", seed=312)
    fd = bvg.Fit(phi, Y, lik, "fp64", "simt", seed=312)
    th = ff.mle()
    st = ff.stats()
    out = {"workload": name, "mle_iters": st["mle_iters"], "mle_term": st["mle_term"], "mle_lp": st["mle_lp"]}
    lp_f, g_f = ff.log_prob(th, False, False)
    lp_d, g_d = fd.log_prob(th, False, False)
    out["lp_fast"], out["lp_fp64"] = lp_f, lp_d
    out["grad_rel_err_fast_vs_fp64"] = float(np.linalg.norm(g_f - g_d) / np.linalg.norm(g_d))
    out["gnorm"] = float(np.linalg.norm(g_d))
    gg = float(g_f @ g_f)
    rows = []
    for a in 10.0 ** np.arange(-14, -3.5, 0.5):
        df = ff.log_prob(th + a * g_f, False, False, grad=False) - lp_f
        dd = fd.log_prob(th + a * g_f, False, False, grad=False) - lp_d
        rows.append({"alpha": float(a), "predicted": a * gg, "fast": df, "fp64": dd})
    out["along_gradient"] = rows
    # noise of the fast evaluator: f at tiny random perturbations (1e-9 relative) of theta
    rng = np.random.default_rng(0)
    v = [ff.log_prob(th * (1 + 1e-9 * rng.normal(size=th.shape)), False, False, grad=False) - lp_f for _ in range(8)]
    w = [fd.log_prob(th * (1 + 1e-9 * rng.normal(size=th.shape)), False, False, grad=False) - lp_d for _ in range(8)]
    out["jitter_fast"], out["jitter_fp64"] = v, w
    print(json.dumps(out))
    ff.close()
    fd.close()


if __name__ == "__main__":
    main()

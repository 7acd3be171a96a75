"""clusterSVGsBayes() at its defaults (R/clusterSVGsBayes.R:63-76: n.clust = 5, n.PCs = 30, fullrank, 30,000 iterations, 300 ELBO
draws every 100 iterations, 1,000 draws) on 1,000 SVGs: the device fit next to the CPU restatement (one thread, 3,000 iterations)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from oracle import oracle as orc  # noqa: E402

rng = np.random.default_rng(312)
N, D, K = 1000, 30, 5
cent = rng.normal(0, 4, (K, D))
lab = rng.integers(0, K, N)
X = cent[lab] + rng.normal(0, 1.0, (N, D))
bvg.engine.gmm_fit(X[:50], 2, "fullrank", n_iter=100, n_draws=10, elbo_samples=10)
out = {}
for alg in ("fullrank", "meanfield"):
    t = time.perf_counter()
    r = bvg.engine.gmm_fit(X, K, alg, tol_rel_obj=0.0)      # tol 0: run all 30,000 iterations (the timing case)
    dt = time.perf_counter() - t
    got = r["soft"].argmax(1)
    purity = sum(np.bincount(got[lab == j], minlength=K).max() for j in range(K)) / N
    out[alg] = dict(wall_s=dt, **r["info"], purity=purity)
    r = bvg.engine.gmm_fit(X, K, alg)                        # the reference's convergence test (tol_rel_obj = 0.01)
    out[alg + "_default_tol"] = dict(**r["info"])
o = orc.GmmOracle(X, K)
t = time.perf_counter()
mu, L, info = o.advi_fullrank(np.zeros(o.D), iter=3000, elbo_samples=300, tol_rel_obj=0.0, eta=out["fullrank"]["eta"], seed=312)
dt = time.perf_counter() - t
out["cpu_port_fullrank"] = dict(wall_s_3000_iters=dt, extrapolated_30000_iters_s=10 * dt, threads=1, iters=info["iters"], elbo=info["elbo"])
out["speedup_fullrank"] = 10 * dt / out["fullrank"]["wall_s"]
print(json.dumps(out))
json.dump(out, open("gpurun_out/r1b_gmm_times.json", "w"), indent=1)

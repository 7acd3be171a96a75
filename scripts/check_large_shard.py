"""One cfg4 shard (50,000 spots x 1,250 genes = what each of 8 GPUs holds at 50k x 10k): parity of the
tcgen05 engine against the fp64 engine and the CPU oracle, kernel timing and HBM roofline at a size
that does not fit in L2 (250 MB)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from bayesvg_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

M, G, k = 50000, 1250, 20
X = bvg.scale_coords(synth.random_points(M, 1))
C = synth.det_kmeans(X, k, 312, iters=30)
ls = bvg.length_scale_kmeans(C)
phi = bvg.basis_build(X, C, ls, "matern", 1.5)
Y, _ = synth.expression(phi, G, "gaussian", 312)
D = 5 + 2 * G + 2 * k + k * G
th = np.random.default_rng(0).normal(0, 0.2, D)
out = {"M": M, "G": G, "k": k}
o = orc.Oracle(phi, Y, "gaussian", nthreads=os.cpu_count())
t0 = time.perf_counter()
rlp, rg = o.log_prob(th, True, True)
out["oracle_eval_s"] = time.perf_counter() - t0
for prec, eng in (("fp64", "simt"), ("fast", "tcgen05")):
    f = bvg.Fit(phi, Y, "gaussian", prec, eng)
    lp, g = f.log_prob(th, True, True)
    out[f"{eng}_lp_rel"] = abs(lp - rlp) / abs(rlp)
    out[f"{eng}_grad_rel"] = float(np.linalg.norm(g - rg) / np.linalg.norm(rg))
    f.set_variational(np.zeros(D))
    f.advi_steps(5, 0.1, reset=True)
    out[f"{eng}_ms_per_eval"] = f.advi_steps(50, 0.1) / 50
    ms = f.time_lik_kernel(20, False, True)
    out[f"{eng}_lik_ms"] = ms
    if eng == "tcgen05":
        b = 4 * M * G + 4 * M * k + 8 * k * G
        out["tcgen05_lik_GBps"] = b / (ms * 1e-3) / 1e9
    f.close()
print(json.dumps(out))
assert out["simt_lp_rel"] < 1e-10 and out["simt_grad_rel"] < 1e-10
assert out["tcgen05_lp_rel"] < 1e-5 and out["tcgen05_grad_rel"] < 1e-4

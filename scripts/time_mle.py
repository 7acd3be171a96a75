"""L-BFGS initialisation at configs[1] (and configs[2] with `nb`): device-resident control flow (lbfgs_dev.cu) against the
host-driven loop (BVG_LBFGS_HOST=1).  Prints seconds, iterations, evaluations, termination and the optimum of both."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bayesvg_b200 as bvg
from bayesvg_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
ds = synth.dataset(name, lambda X, C, ls, kern, nu: bvg.basis_build(X, C, ls, kern, nu))
out = {}
for label, host in (("host", "1"), ("device", "0"), ("host2", "1"), ("device2", "0")):
    os.environ["BVG_LBFGS_HOST"] = host
    f = bvg.Fit(ds["phi"], ds["Y"], ds["likelihood"], "fast", "auto", seed=312)
    t0 = time.perf_counter()
    th = f.mle()
    dt = time.perf_counter() - t0
    st = f.stats()
    out[label] = dict(sec=dt, iters=st["mle_iters"], fevals=st["mle_fevals"], term=st["mle_term"], lp=st["mle_lp"],
                      fp64_from=st["mle_fp64_from"], us_per_iter=1e6 * dt / max(st["mle_iters"], 1), th_norm=float(np.linalg.norm(th)))
    f.close()
print(json.dumps(out, indent=1))

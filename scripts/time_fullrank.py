"""algorithm = "fullrank" at BASELINE config 2 (4,992 spots x 3,000 genes, k = 20: D = 66,045, L and its step-size
history = 34.9 GB in HBM): device time per ELBO-gradient evaluation + step, per ELBO estimate, and HBM throughput of the
fused update pass (32 bytes per entry of the packed triangle)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from bayesvg_b200 import synth  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
X = synth.scale_coords(synth.visium_grid())
Cc = synth.det_kmeans(X, 20, 312, iters=30)
phi = bvg.basis_build(X, Cc, bvg.length_scale_kmeans(Cc), "matern", 2.5)
Y, _ = synth.expression(phi, G, "gaussian", 312)
f = bvg.Fit(phi, Y, "gaussian", "fast", "auto", algorithm="fullrank", seed=312)
D = f.D
f.set_variational(np.zeros(D))
f.advi_steps(3, 0.1, reset=True)
ms = f.advi_steps(20, 0.1) / 20
t = time.perf_counter()
e, ms_elbo = f.elbo(150)
n = D * (D + 1) // 2
out = dict(workload=f"4,992 x {G}, k = 20, gaussian, fullrank", D=D, tri_entries=n, state_GB=16 * n / 1e9,
           ms_per_eval_and_step=ms, update_pass_GBps=32 * n / (ms * 1e-3) / 1e9, ms_per_elbo_150=ms_elbo, elbo=e)
t = time.perf_counter()
f.summarize()
out["summary_s"] = time.perf_counter() - t
print(json.dumps(out))
json.dump(out, open("gpurun_out/r1b_fullrank_cfg2.json", "w"), indent=1)
f.close()

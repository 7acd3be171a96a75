"""algorithm = "pathfinder" at BASELINE config 2 (4,992 spots x 3,000 genes, k = 20): wall-clock of the reference's call
(:366-375: num_paths = n.cores = 2, max_lbfgs_iters = 200, num_elbo_draws = 50, history_size = 25, draws = 1000)."""
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
Y, amp_true = synth.expression(phi, G, "gaussian", 312)
out = {}
for alg, kw in (("pathfinder", dict(elbo_samples=50, pf_num_paths=2)), ("meanfield", {})):
    f = bvg.Fit(phi, Y, "gaussian", "fast", "auto", algorithm=alg, seed=312, **kw)
    t = time.perf_counter()
    tab = f.run()
    dt = time.perf_counter() - t
    st = f.stats()
    out[alg] = dict(wall_s=dt, iters=st["advi_iters"], elbo=st["elbo"], eta_or_khat=st["eta"], launches=st["kernel_launches"],
                    trace=f.trace().tolist()[:4])
    out[alg + "_mean"] = tab[:, 0]
    f.close()
from scipy.stats import spearmanr  # noqa: E402
pm, mm = out.pop("pathfinder_mean"), out.pop("meanfield_mean")
out["spearman_pathfinder_vs_meanfield"] = float(spearmanr(pm, mm).correlation)
out["spearman_pathfinder_vs_truth"] = float(spearmanr(pm, amp_true).correlation)
out["spearman_meanfield_vs_truth"] = float(spearmanr(mm, amp_true).correlation)
top = 1000 if G >= 3000 else G // 3
out["top_overlap_pathfinder_vs_meanfield"] = len(set(np.argsort(-pm)[:top]) & set(np.argsort(-mm)[:top])) / top
print(json.dumps(out))
json.dump(out, open("gpurun_out/r1b_pathfinder_cfg2.json", "w"), indent=1)

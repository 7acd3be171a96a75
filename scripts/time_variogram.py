"""lscale.estimator = "variogram" on the device at BASELINE shapes: wall-clock of bvg_variogram_lscale."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from bayesvg_b200 import synth  # noqa: E402

out = []
for name, coords, G in (("cfg1", synth.masked_blob(synth.visium_grid(), 2444), 3000), ("cfg2", synth.visium_grid(), 3000)):
    X = bvg.scale_coords(coords)
    C = synth.det_kmeans(X, 20, 312, iters=30)
    phi = bvg.basis_build(X, C, bvg.length_scale_kmeans(C), "matern", 1.5)
    Y, _ = synth.expression(phi, G, "gaussian", 312)
    bvg.engine.variogram_lscale(X[:200], Y[:200, :5].T, 1.5)
    t = time.perf_counter()
    ls, r = bvg.engine.variogram_lscale(X, np.ascontiguousarray(Y.T), 1.5, return_ranges=True)
    out.append(dict(workload=name, M=X.shape[0], G=G, wall_s=round(time.perf_counter() - t, 4), lscale=ls, lscale_kmeans=bvg.length_scale_kmeans(C),
                    n_failed=int(np.isnan(r).sum()), pair_gene_ops=float(X.shape[0]) ** 2 / 2 * G))
    print(json.dumps(out[-1]))
json.dump(out, open("gpurun_out/r1b_variogram_times.json", "w"), indent=1)

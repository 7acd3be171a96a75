"""Wall-clock of bvg_kmeans (R/findSpatiallyVariableFeaturesBayes.R:200-204 on the device) at the BASELINE shapes."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bayesvg_b200 as bvg  # noqa: E402

out = []
for name, M, k in (("cfg2", 4992, 20), ("cfg4", 50000, 20), ("cfg5", 200000, 400)):
    X = bvg.scale_coords(np.random.default_rng(1).uniform(0, 1, (M, 2)))
    bvg.engine.kmeans(X[:1000], 5, 5, 1)
    t = time.perf_counter()
    C, info = bvg.engine.kmeans(X, k, 100, 10, seed=312, return_info=True)
    out.append(dict(workload=name, M=M, k=k, iter_max=100, nstart=10, wall_s=round(time.perf_counter() - t, 4), **info))
    print(json.dumps(out[-1]))
json.dump(out, open("gpurun_out/r1_kmeans_times.json", "w"), indent=1)

"""python scripts/profile_chain.py [n_steps] -- a short ELBO-gradient chain at cfg2 (4,992 x 3,000, k = 20) for ncu:
   ncu --set full --clock-control none --import-source on -k regex:k_gene -s 20 -c 2 -o gpurun_out/prof python scripts/profile_chain.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from bayesvg_b200 import synth  # noqa: E402
from bayesvg_b200.api import length_scale_kmeans  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lik = sys.argv[2] if len(sys.argv) > 2 else "gaussian"
X = synth.scale_coords(synth.visium_grid())
Cc = synth.det_kmeans(X, 20, 312, iters=30)
phi = bvg.basis_build(X, Cc, length_scale_kmeans(Cc), "matern", 2.5)
Y, _ = synth.expression(phi, 3000, lik, 312)
f = bvg.Fit(phi, Y, lik, "fast", "auto", seed=312)
f.set_variational(np.zeros(f.D), None)
ms = f.advi_steps(n, 0.1, reset=True)
print(f"{n} ELBO-gradient evaluations in {ms:.3f} ms")
f.close()

"""python scripts/profile_r2.py [gaussian|nb] -- one of each hot launch of round 2 at cfg2 for ncu: a 40-iteration stretch of
k_advi_persist, one ELBO estimate (k_prep batch + k_elbo_persist, 150 draws), and backward / forward launches of k_lik_tc."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from bayesvg_b200 import synth  # noqa: E402
from bayesvg_b200.api import length_scale_kmeans  # noqa: E402

lik = sys.argv[1] if len(sys.argv) > 1 else "gaussian"
X = synth.scale_coords(synth.visium_grid())
Cc = synth.det_kmeans(X, 20, 312, iters=30)
phi = bvg.basis_build(X, Cc, length_scale_kmeans(Cc), "matern", 2.5)
Y, _ = synth.expression(phi, 3000, lik, 312)
f = bvg.Fit(phi, Y, lik, "fast", "tcgen05", seed=312, elbo_suffstat=0)   # the dense ELBO kernel is what is profiled
f.set_variational(np.zeros(f.D), None)
for _ in range(3):
    ms = f.advi_steps(40, 0.1, reset=True)
e, ms_e = f.elbo(150)
e, ms_e = f.elbo(150)
print(f"40 ELBO-gradient evaluations {ms:.3f} ms; ELBO(150 draws) {ms_e:.3f} ms; k_lik_tc cold {1e3 * f.time_lik_kernel(5, True, True):.1f} us")
f.close()

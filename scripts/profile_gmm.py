"""a short clusterSVGsBayes() fit at the default shape for ncu: ncu --set full -k regex:k_gmm_elbo -c 1 python scripts/profile_gmm.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402

rng = np.random.default_rng(312)
N, D, K = 1000, 30, 5
cent = rng.normal(0, 4, (K, D))
X = cent[rng.integers(0, K, N)] + rng.normal(0, 1.0, (N, D))
alg = sys.argv[1] if len(sys.argv) > 1 else "meanfield"
r = bvg.engine.gmm_fit(X, K, alg, n_iter=300, n_draws=50, eta=0.1)
print(r["info"])

"""ncu --set full -k regex:k_vg_gamma -c 1 python scripts/profile_variogram.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from bayesvg_b200 import synth  # noqa: E402

X = bvg.scale_coords(synth.visium_grid())
Z = np.random.default_rng(0).normal(size=(1024, X.shape[0]))
print(bvg.engine.variogram_lscale(X, Z, 1.5))

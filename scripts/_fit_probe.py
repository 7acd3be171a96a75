import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import bayesvg_b200 as bvg
from bayesvg_b200 import synth
os.environ["BVG_TIMING"] = "1"
ds = synth.dataset("cfg2", lambda X, C, ls, kern, nu: bvg.basis_build(X, C, ls, kern, nu))
Y = ds["Y"]
print("Y flags", Y.flags["F_CONTIGUOUS"], Y.dtype, Y.shape)
for rep in range(3):
    t0 = time.perf_counter()
    f = bvg.Fit(ds["phi"], Y, "gaussian", "fast", "auto", seed=312)
    t1 = time.perf_counter()
    f.run()
    t2 = time.perf_counter()
    s = f.stats()
    print(f"rep {rep}: create+set {1e3*(t1-t0):.1f} ms, run {1e3*(t2-t1):.1f} ms; h2d {1e3*s['sec_h2d']:.1f} mle {1e3*s['sec_mle']:.1f} adapt {1e3*s['sec_adapt']:.1f} sga {1e3*s['sec_sga']:.1f} draws {1e3*s['sec_draws']:.1f} total {1e3*s['sec_total']:.1f}")
    f.close()

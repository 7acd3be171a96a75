import os, sys, time
sys.path.insert(0, "/root/repo")
import bayesvg_b200 as bvg
from bayesvg_b200 import synth
ds = synth.dataset("cfg2", lambda X, C, ls, kern, nu: bvg.basis_build(X, C, ls, kern, nu))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    f = bvg.Fit(ds["phi"], ds["Y"], "gaussian", "fast", "auto", seed=312, mle_iter=n)
    t0 = time.perf_counter(); f.mle(); dt = time.perf_counter() - t0
    print(f.stats()["mle_iters"], f.stats()["mle_fevals"], f"{dt*1e3:.1f} ms")
    f.close()

"""python scripts/stage_times.py [M G k] -- device time of the stages of one ELBO-gradient evaluation (CUDA
events on the engine's stream) and of the evaluation chain as the ADVI loop runs it."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from bayesvg_b200 import synth  # noqa: E402
from bayesvg_b200.api import length_scale_kmeans  # noqa: E402


def main():
    M, G, k = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4992, 3000, 20)
    lik = sys.argv[4] if len(sys.argv) > 4 else "gaussian"
    X = synth.scale_coords(synth.visium_grid() if M == 4992 else synth.random_points(M, 1))
    Cc = synth.det_kmeans(X, k, 312, iters=30) if k <= 30 else bvg.kmeans_centers(X, k, 30, 1, rng=312)
    phi = bvg.basis_build(X, Cc, length_scale_kmeans(Cc), "matern", 2.5)
    Y, _ = synth.expression(phi, G, lik, 312)
    f = bvg.Fit(phi, Y, lik, "fast", "auto", seed=312)
    f.set_variational(np.zeros(f.D), None)
    f.advi_steps(20, 0.1, reset=True)
    st = f.stage_times(50)
    ms = min(f.advi_steps(200, 0.1) for _ in range(5)) / 200
    print(f"M={M} G={G} k={k} {lik}: prep {1e3 * st[0]:.1f} us, likelihood {1e3 * st[1]:.1f} us, gene+finish {1e3 * st[2]:.1f} us "
          f"(each stage timed alone); chained evaluation {1e3 * ms:.1f} us = {1e3 / ms:.0f} evals/s; "
          f"likelihood kernel alone back-to-back {1e3 * f.time_lik_kernel(50, False, True):.1f} us, cold L2 {1e3 * f.time_lik_kernel(20, True, True):.1f} us")
    if os.environ.get("BVG_TC_TRACE"):
        for rep in range(2):
            t = f.tc_trace(31)
            t0 = t[30, 0]
            print("tc trace (cycles since kernel start): setup done", t[30, 1] - t0, " before final sync", t[30, 2] - t0, " end", t[30, 3] - t0)
            for c in range(14):
                if t[c, 0]:
                    print(f"  chunk {c:2d}: " + " ".join(f"{(x - t0) if x else 0:6d}" for x in t[c]))
    if os.environ.get("BVG_TRACE"):
        for mode, fuse, ah, what in ((3, 0, 0, "lp only, no finish"), (3, 3, 0, "lp only + finish"), (0, 0, 0, "grad out, no finish"),
                                     (1, 0, 0, "update, no finish"), (1, 3, 0, "update + finish"), (1, 0, 1, "update + draw-ahead, no finish"),
                                     (1, 3, 1, "update + draw-ahead + finish")):
            print(f"k_gene x200 back to back, {what}: {1e3 * f.time_gene(200, mode, fuse, ah):.2f} us")
        for ah in (0, 2):
            for _ in range(3):
                print(f"k_gene trace (ahead={ah}) ns:", f.gene_trace(ah)[:16])
    f.close()


if __name__ == "__main__":
    main()

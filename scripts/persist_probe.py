"""python scripts/persist_probe.py [M G k lik] -- the persistent ADVI kernel (k_advi_persist) against the two-launch
evaluation chain on the same problem: device time per ELBO-gradient evaluation (CUDA events on the engine's stream)
and the in-kernel phase breakdown of CTA 0 (clock64 stamps, bvg_debug_persist_trace)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from bayesvg_b200 import synth  # noqa: E402
from bayesvg_b200.api import length_scale_kmeans  # noqa: E402


def make_fit(phi, Y, lik, no_persist):
    os.environ["BVG_NO_PERSIST"] = "1" if no_persist else "0"
    f = bvg.Fit(phi, Y, lik, "fast", "tcgen05", seed=312)
    f.set_variational(np.zeros(f.D), None)
    f.advi_steps(20, 0.1, reset=True)
    return f


def main():
    M, G, k = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4992, 3000, 20)
    lik = sys.argv[4] if len(sys.argv) > 4 else "gaussian"
    X = synth.scale_coords(synth.visium_grid() if M == 4992 else synth.random_points(M, 1))
    Cc = synth.det_kmeans(X, k, 312, iters=30)
    phi = bvg.basis_build(X, Cc, length_scale_kmeans(Cc), "matern", 2.5)
    Y, _ = synth.expression(phi, G, lik, 312)
    out = {"M": M, "G": G, "k": k, "likelihood": lik}
    out["pre"] = os.environ.get("BVG_PERSIST_PRE", "default")
    for name, nop in ((("two_launch", True),) if not os.environ.get("PROBE_SKIP2") else ()) + (("persistent", False),):
        f = make_fit(phi, Y, lik, nop)
        ms = min(f.advi_steps(200, 0.1) for _ in range(5)) / 200
        out[name + "_us_per_eval"] = 1e3 * ms
        if not nop:
            out["stretch_of_1_us"] = 1e3 * min(f.advi_steps(1, 0.1) for _ in range(5))
            t, c = f.persist_trace(24)
            t = t[:24]
            out["gene_phase_fine_cycles_median"] = {n: float(np.median(t[4:, b] - t[4:, a_])) for n, a_, b in (
                ("red_to_geom", 1, 8), ("load0", 8, 9), ("load1", 9, 6), ("tile_wait", 6, 7), ("process0", 7, 10), ("process1", 10, 11),
                ("to_row_written", 11, 12), ("sync", 12, 13), ("row_sum_store", 13, 14), ("red", 14, 15), ("grid_wait", 15, 3))}
            d = np.diff(t[:, :6], axis=1)[4:]          # skip the first iterations (cold rings)
            period = np.diff(t[4:, 0])
            out["normals_then_tile_wait_cycles_median"] = [float(np.median(t[4:, 6] - t[4:, 1])), float(np.median(t[4:, 7] - t[4:, 6]))]
            t0 = t[8, 0]
            out["pass_of_iteration_8"] = {"coeff_stored": int(c[31, 0] - t0), "chunk_loop_done": int(c[31, 1] - t0), "d2_read": int(c[31, 2] - t0),
                                          "partials_written": int(c[31, 3] - t0), "pass_end_stamp": int(t[8, 1] - t0),
                                          "chunks": [[int(v - t0) if v else 0 for v in row] for row in c[:14] if row[0]]}
            out["phase_cycles_median"] = {n: float(np.median(d[:, i])) for i, n in
                                          enumerate(["passes", "gene_phases", "grid_barrier", "reduce_exchange", "shared_update"])}
            out["iteration_cycles_median"] = float(np.median(period))
        f.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

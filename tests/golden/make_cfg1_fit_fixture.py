"""python tests/golden/make_cfg1_fit_fixture.py [threads] -- the whole default fit of the CPU oracle (L-BFGS init, eta adaptation,
3,000 SGA iterations, 1,000 draws, amplitude summaries) on the stand-in for BASELINE.json configs[0] (seu_brain: 2,444 in-tissue
spots x 3,000 HVGs, k = 20 Matern-2.5, gaussian; bayesvg_b200.synth.dataset("cfg1")), stored as tests/golden/cfg1_fit.npz.
tests/test_gpu_parity.py::test_cfg1_fit_agrees_with_the_oracle_fit_fixture compares the CUDA fast path's summaries with it
(north_star's second check: Spearman and top-1,000 overlap).  Takes ~1-2 minutes of CPU on 16 threads."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bayesvg_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    nthreads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    ds = synth.dataset("cfg1", lambda X, C, ls, kern, nu: orc.basis(X, C, ls, kern, nu))
    t0 = time.perf_counter()
    o = orc.Oracle(ds["phi"], ds["Y"], "gaussian", nthreads=nthreads)
    th, li = o.lbfgs(max_iter=1000)
    mu, om, info = o.advi(th, iter=3000, elbo_samples=150, seed=312)
    tab = o.summary(mu, om, 1000, 312)
    wall = time.perf_counter() - t0
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cfg1_fit.npz"), amplitude_mean=tab[:, 0], amplitude_median=tab[:, 1],
                        amplitude_sd=tab[:, 2], eta=info["eta"], advi_iters=info["iters"], wall_s=wall, threads=nthreads,
                        y_checksum=float(np.abs(ds["Y"]).sum()), phi_checksum=float(np.abs(ds["phi"]).sum()), planted_amplitude=ds["amp"])
    print(f"cfg1 oracle fit: {wall:.1f} s on {nthreads} threads, eta {info['eta']}, {info['iters']} iterations")


if __name__ == "__main__":
    main()

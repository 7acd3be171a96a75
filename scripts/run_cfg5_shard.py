"""One BASELINE config-5 shard on one B200: 200,000 spots x 2,500 genes (what each of 8 GPUs holds at 200,000 x 20,000),
k = 400 basis functions, periodic kernel, gaussian, mean-field ADVI -- the whole path through the public API:
device k-means (bvg_kmeans) -> basis (bvg_basis_build) -> fit (bvg_fit_run, large-basis tcgen05 kernels) -> summaries."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from bayesvg_b200 import synth  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
k = int(sys.argv[3]) if len(sys.argv) > 3 else 400
out = dict(M=M, G=G, k=k)
t = time.perf_counter()
X = bvg.scale_coords(synth.random_points(M, 2))
C = bvg.kmeans_centers(X, k, 100, 10, rng=312)
ls = bvg.length_scale_kmeans(C)
out["kmeans_s"] = time.perf_counter() - t
# kernel.period ~ 2 x the scaled median nearest-centroid distance (keeps the periodic kernel matrix full rank)
d = np.sqrt(((C[:, None, :] - C[None]) ** 2).sum(-1))
np.fill_diagonal(d, np.inf)
period = 2.0 * float(np.median(d.min(1)))
out["lscale"], out["period"] = ls, period
t = time.perf_counter()
phi = bvg.basis_build(X, C, ls, "periodic", 1.5, period)
out["basis_s"] = time.perf_counter() - t
out["basis_orth_err"] = float(np.abs(phi[:, :40].T @ phi[:, :40] - np.eye(40)).max())
t = time.perf_counter()
rng = np.random.default_rng(312)
amp = np.exp(rng.normal(-1, 1, G))
amp[rng.random(G) < 0.3] = 0
Y = np.empty((M, G), order="F", dtype=np.float32)
for g0 in range(0, G, 250):   # in blocks: the dense fp64 product would need 4 GB at once
    g1 = min(G, g0 + 250)
    al = rng.normal(size=(k, g1 - g0))
    Y[:, g0:g1] = (rng.normal(0, 0.5, g1 - g0)[None] + (amp[g0:g1] ** 2)[None] * (phi @ al) * np.sqrt(M / k) + rng.normal(size=(M, g1 - g0)))
mu, sd = float(Y.mean(dtype=np.float64)), float(Y.std(dtype=np.float64, ddof=1))
Y -= mu
Y /= sd
out["synth_s"] = time.perf_counter() - t
t = time.perf_counter()
f = bvg.Fit(phi, Y, "gaussian", "fast", "auto", seed=312, verbose=1)
out["create_and_h2d_s"] = time.perf_counter() - t
f.set_variational(np.zeros(f.D))
f.advi_steps(3, 0.1, reset=True)
out["ms_per_elbo_grad_eval"] = f.advi_steps(20, 0.1) / 20
out["ms_lik_pass_alone"] = f.time_lik_kernel(10, False, True)
flops = 4.0 * k * M * G + 12.0 * M * G
out["algorithmic_tflops_eval"] = flops / (out["ms_per_elbo_grad_eval"] * 1e-3) / 1e12
out["executed_bf16_tflops_lik_pass"] = 3 * 4.0 * (k + 1) * M * G / (out["ms_lik_pass_alone"] * 1e-3) / 1e12
t = time.perf_counter()
tab = f.run()
out["fit_wall_s"] = time.perf_counter() - t
st = f.stats()
out["fit"] = {q: st[q] for q in ("advi_iters", "grad_evals", "elbo_evals", "eta", "elbo", "mle_iters", "sec_mle", "sec_adapt", "sec_sga", "sec_draws", "converged")}
from scipy.stats import spearmanr  # noqa: E402
out["spearman_amplitude_vs_truth"] = float(spearmanr(tab[:, 0], amp).correlation)
f.close()
print(json.dumps(out))
json.dump(out, open("gpurun_out/r1b_cfg5_shard.json", "w"), indent=1)

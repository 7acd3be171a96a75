"""torchrun --nproc-per-node N scripts/persist_probe_multi.py -- phase breakdown (clock64 stamps of CTA 0, every rank) of the
persistent ADVI kernel on a gene-sharded fit: cfg2 per GPU (weak) or one cfg4 shard per GPU (strong); BVG_XCHG_ALL=0|1 picks
the exchange form."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesvg_b200 as bvg  # noqa: E402
from bayesvg_b200 import dist, synth  # noqa: E402
from bayesvg_b200.api import length_scale_kmeans  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg4 = len(sys.argv) > 1 and sys.argv[1] == "cfg4"
    M, Gl = (50000, 10000 // world) if cfg4 else (4992, 3000)
    X = synth.scale_coords(synth.random_points(M, 1) if cfg4 else synth.visium_grid())
    C = synth.det_kmeans(X, 20, 312, iters=30)
    phi = bvg.basis_build(X, C, length_scale_kmeans(C), "matern", 1.5 if cfg4 else 2.5, device=local)
    Y, _ = synth.expression(phi, Gl, "gaussian", 312 + rank)
    f = dist.make_sharded_fit(phi, Y, dist.Shard(rank, world), Gl * world, Gl * rank, "gaussian", "fast", "tcgen05", dict(device=local, seed=312), "peer")
    f.set_variational(np.zeros(f.D), None)
    f.advi_steps(20, 0.1, reset=True)
    ms = min(f.advi_steps(200, 0.1) for _ in range(5)) / 200
    t, _ = f.persist_trace(24)
    t = t[:24]
    d = np.diff(t[:, :6], axis=1)[4:]
    out = {"rank": rank, "us_per_eval": 1e3 * ms, "xchg_all": os.environ.get("BVG_XCHG_ALL", "default"),
           "phase_cycles_median": {n: float(np.median(d[:, i])) for i, n in enumerate(["passes", "gene_phases", "grid_barrier", "reduce_exchange", "shared_update"])},
           "iteration_cycles_median": float(np.median(np.diff(t[4:, 0])))}
    f.close()
    for r in range(world):
        td.barrier()
        if r == rank and rank in (0, world - 1):
            print(json.dumps(out), flush=True)
    td.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- ELBO-gradient evaluations per second of the bayesVG SVG hot path on B200.

Workload (BASELINE.json configs[1]): synthetic full Visium grid, 4,992 spots x 3,000 genes, k = 20
Matern-2.5 basis, Gaussian likelihood, mean-field ADVI.  One STEP = --evals-per-step (100) consecutive
ELBO-gradient evaluations = the stretch of the ADVI loop between two convergence checks; one
evaluation = one Philox draw of all D = 66,045 parameters, the fused likelihood pass over all
spot x gene pairs (forward + analytic backward) and the mean-field update.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 (torchrun, one rank per GPU): weak scaling -- every rank holds its own 3,000 genes (G_total =
3,000 N), the 2k+10 cross-gene sums are all-reduced inside every evaluation (--comm peer: fused into the
per-gene kernel over peer memory; --comm nccl: ncclAllReduce between kernels), and `value` counts
3,000-gene evaluations (N per global evaluation).

Every line also carries `strong_cfg4` (BASELINE.json configs[3], the split north_star names: 50,000 spots x 10,000 genes in
total, gene-sharded over the N ranks -- STRONG scaling: value = evaluations of the whole problem per second, plus the wall of
the whole default fit), so the driver's 1/2/4/8-GPU run records it next to the weak-scaling line; for N > 1 also
`parity_check` (a gene-sharded problem against its unsharded twin on the same GPUs, before anything is timed).  At N = 1 the
line further carries `cfg3_nb` (configs[2]), `fp64_engine`, and `cfg1_fit` (the stand-in for configs[0] with a bounded CPU leg).

--impl reference times the reference's CPU path.  The reference is R + CmdStan, neither of which
exists in this image, so that arm is the fp64 C restatement in oracle/ ("port") running
the same evaluation with all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_SPOTS, G_GENES, K_BASIS = 4992, 3000, 20
# --workload cfg4 (BASELINE.json configs[3]): 50,000 spots x 10,000 genes, Matern-1.5, gene-sharded over the ranks
# (STRONG scaling: the 10,000 genes are split, `value` counts evaluations of the whole problem)
CFG4_M, CFG4_G = 50000, 10000


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = index, [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def make_dataset(basis_fn, seed, workload="cfg2", n_genes=G_GENES):
    from bayesvg_b200 import synth
    from bayesvg_b200.api import length_scale_kmeans
    if workload == "cfg4":
        X = synth.scale_coords(synth.random_points(CFG4_M, 1))
        C = synth.det_kmeans(X, K_BASIS, 312, iters=30)
    else:
        X = synth.scale_coords(synth.visium_grid())
        C = synth.det_kmeans(X, K_BASIS, 312)
    ls = length_scale_kmeans(C)
    phi = basis_fn(X, C, ls)
    Y, _ = synth.expression(phi, n_genes, "gaussian", seed)
    return phi, Y


def st_engine(fit):
    return {1: "simt", 2: "tcgen05"}.get(fit.stats()["engine_used"], "?")


def cpu_evals(phi, Y, min_seconds, min_evals, nthreads):
    """time ELBO-gradient evaluations (draw + log_prob_grad + meanfield update) of the CPU restatement"""
    from oracle import oracle as orc
    o = orc.Oracle(phi, Y, "gaussian", nthreads=nthreads)
    mu, om, hist = np.zeros(o.D), np.zeros(o.D), np.zeros(2 * o.D)
    o.advi_step(mu, om, hist, 0.1, 1, 312, 0)  # warm-up (page in Y)
    n, t0 = 0, time.perf_counter()
    while n < min_evals or time.perf_counter() - t0 < min_seconds:
        o.advi_step(mu, om, hist, 0.1, n + 2, 312, n + 1)
        n += 1
    return n, time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as orc
    nthreads = os.cpu_count() or 1
    phi, Y = make_dataset(lambda X, C, ls: orc.basis(X, C, ls, "matern", 2.5), 312)
    o = orc.Oracle(phi, Y, "gaussian", nthreads=nthreads)
    mu, om, hist = np.zeros(o.D), np.zeros(o.D), np.zeros(2 * o.D)
    per_step = args.ref_evals_per_step
    t = 0
    for _ in range(args.warmup):
        for _ in range(per_step):
            o.advi_step(mu, om, hist, 0.1, t + 1, 312, t)
            t += 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(per_step):
            o.advi_step(mu, om, hist, 0.1, t + 1, 312, t)
            t += 1
    dt = time.perf_counter() - t0
    value = args.steps * per_step / dt
    sample = f"{args.steps} steps x {per_step} ELBO-gradient evals of the full 4,992 x 3,000 workload (fp64, OpenMP x {nthreads})"
    print(json.dumps({
        "impl": "reference", "metric": "elbo_grad_evals_per_sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: 4,992 spots x 3,000 genes, k=20 Matern-2.5, gaussian, meanfield ADVI",
                   "evals_per_step": per_step,
                   "note": "R/CmdStan are absent from the image: this is the fp64 C restatement (oracle/) of the "
                           "reference's log-density/gradient + ADVI update; it has no AD tape, no JSON/CSV, no g++ "
                           "compile, so it is faster than the real CmdStan path"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _max_over_ranks(x, world):
    if world == 1:
        return float(x)
    import torch
    import torch.distributed as td
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def parity_check(rank, world, local_rank, engine, comm):
    """N > 1: a gene-sharded problem on the N ranks against its unsharded twin (every rank holds both): log density,
    gradient, a 25-step ADVI trajectory through the persistent kernel's in-kernel exchange, and an ELBO estimate."""
    import torch
    import torch.distributed as td

    import bayesvg_b200 as bvg
    from bayesvg_b200 import dist as bdist
    from bayesvg_b200 import synth
    from bayesvg_b200.api import length_scale_kmeans
    M, G, k, lik = 1500, 128 * world + 37, 20, "gaussian"
    X = synth.scale_coords(synth.random_points(M, 7))
    C = synth.det_kmeans(X, k, 312, iters=30)
    phi = bvg.basis_build(X, C, length_scale_kmeans(C), "matern", 2.5, device=local_rank)
    Y, _ = synth.expression(phi, G, lik, 99)
    sh = bdist.Shard(rank, world)
    g0, g1 = sh.gene_range(G)
    idx = bdist.shard_index(lik, G, k, g0, g1)
    cfg = dict(device=local_rank, seed=312)
    fs = bdist.make_sharded_fit(phi, np.asfortranarray(Y[:, g0:g1]), sh, G, g0, lik, "fast", engine, cfg, comm)
    ff = bvg.Fit(phi, Y, lik, "fast", engine, **cfg)
    theta = np.random.default_rng(1).normal(0, 0.3, ff.D)
    lp_s, g_s = fs.log_prob(theta[idx], True, False)
    lp_f, g_f = ff.log_prob(theta, True, False)
    e_lp = abs(lp_s - lp_f) / abs(lp_f)
    e_g = float(np.linalg.norm(g_s - g_f[idx]) / np.linalg.norm(g_f[idx]))
    for f, th in ((fs, theta[idx]), (ff, theta)):
        f.set_variational(0.1 * th)
        f.advi_steps(0, 0.1, reset=True)
        f.advi_steps(25, 0.1)
    mu_s, om_s = fs.get_variational()
    mu_f, om_f = ff.get_variational()
    e_t = float(max(np.linalg.norm(mu_s - mu_f[idx]) / np.linalg.norm(mu_f[idx]), np.linalg.norm(om_s - om_f[idx]) / np.linalg.norm(om_f[idx])))
    el_s, el_f = fs.elbo(20)[0], ff.elbo(20)[0]
    e_e = abs(el_s - el_f) / abs(el_f)
    fs.close()
    ff.close()
    worst = torch.tensor([e_lp, e_g, e_t, e_e], device="cuda", dtype=torch.float64)
    td.all_reduce(worst, op=td.ReduceOp.MAX)
    e_lp, e_g, e_t, e_e = (float(v) for v in worst.tolist())
    # fast path (BF16x3): the shards sum the same products in a different order; gates as tests/test_gpu_multi.py
    return {"what": f"{M} spots x {G} genes, k={k}, gaussian, fast/{engine}, gene-sharded over {world} ranks vs unsharded on every rank",
            "log_prob_rel": e_lp, "grad_rel": e_g, "advi_25_steps_rel": e_t, "elbo_rel": e_e,
            "ok": bool(e_lp < 2e-5 and e_g < 2e-5 and e_t < 2e-2 and e_e < 2e-5)}


def strong_cfg4(args, rank, world, local_rank, barrier):
    """BASELINE.json configs[3]: 50,000 spots x 10,000 genes, Matern-1.5, the genes split over the ranks (strong scaling)."""
    import bayesvg_b200 as bvg
    from bayesvg_b200 import dist as bdist
    g_lo, g_hi = bdist.Shard(rank, world).gene_range(CFG4_G)
    t0 = time.perf_counter()
    phi, Y = make_dataset(lambda X, C, ls: bvg.basis_build(X, C, ls, "matern", 1.5, device=local_rank), 312 + rank, "cfg4", g_hi - g_lo)
    t_data = time.perf_counter() - t0
    cfg = dict(device=local_rank, seed=312)

    def make():
        if world > 1:
            return bdist.make_sharded_fit(phi, Y, bdist.Shard(rank, world), CFG4_G, g_lo, "gaussian", "fast", args.engine, cfg, args.comm)
        return bvg.Fit(phi, Y, "gaussian", "fast", args.engine, **cfg)

    fit = make()
    E = args.evals_per_step
    fit.set_variational(np.zeros(fit.D), None)
    fit.advi_steps(0, 0.1, reset=True)
    for _ in range(3):
        fit.advi_steps(E, 0.1)
    steps = max(2, min(args.steps, 5))
    barrier()
    ms = 0.0
    for _ in range(steps):
        ms += fit.advi_steps(E, 0.1)   # Y of a shard (>= 250 MB) exceeds L2: every evaluation streams it from HBM
    barrier()
    ms = _max_over_ranks(ms, world)
    ms_lik = fit.time_lik_kernel(10, False, True)
    fit.close()
    barrier()
    t0 = time.perf_counter()
    f3 = make()
    f3.run()
    barrier()
    wall = _max_over_ranks(time.perf_counter() - t0, world)
    s3 = f3.stats()
    f3.close()
    peak, _ = load_peaks()
    Gl = g_hi - g_lo
    bl = 4 * CFG4_M * Gl + 4 * CFG4_M * K_BASIS + 8 * K_BASIS * Gl
    return {"workload": "cfg4: 50,000 spots x 10,000 genes in total, k=20 Matern-1.5, gaussian, meanfield ADVI; genes split over the ranks",
            "scaling": "strong", "value": steps * E / (ms * 1e-3), "unit": "evals/s (whole 10,000-gene problem)", "n_gpus": world,
            "us_per_eval": 1e3 * ms / (steps * E), "steps": steps, "evals_per_step": E, "genes_per_gpu": Gl,
            "likelihood_kernel_us": 1e3 * ms_lik, "likelihood_kernel_hbm_frac": bl / (ms_lik * 1e-3) / 1e9 / peak,
            "fit": {"wall_s": wall, "grad_evals": s3["grad_evals"], "elbo_evals": s3["elbo_evals"], "advi_iters": s3["advi_iters"],
                    "mle_iters": s3["mle_iters"], "mle_term": s3["mle_term"], "sec_mle": s3["sec_mle"], "sec_adapt": s3["sec_adapt"],
                    "sec_sga": s3["sec_sga"], "sec_draws": s3["sec_draws"], "converged": s3["converged"], "elbo": s3["elbo"]},
            "synthetic_data_seconds": t_data}


def extra_blocks_n1(args, local_rank, phi_cfg2, Y_cfg2):
    """N = 1 only: BASELINE configs[2] (nb, exp_quad), the fp64 engine on configs[1], and the configs[0] stand-in fit with a
    bounded CPU leg, so that those numbers are in the driver's record and not only in profiles/."""
    import bayesvg_b200 as bvg
    from bayesvg_b200 import synth
    from bayesvg_b200.api import kmeans_centers, length_scale_kmeans
    peak, _ = load_peaks()
    cfg = dict(device=local_rank, seed=312)
    out = {}
    # ---- cfg3: counts, nb likelihood, exp_quad kernel
    X = synth.scale_coords(synth.visium_grid())
    C = synth.det_kmeans(X, K_BASIS, 312)
    phi3 = bvg.basis_build(X, C, length_scale_kmeans(C), "exp_quad", 0.0, device=local_rank)
    Y3, _ = synth.expression(phi3, G_GENES, "nb", 312)
    Y3 = np.asfortranarray(Y3.astype(np.int32))
    f = bvg.Fit(phi3, Y3, "nb", "fast", args.engine, **cfg)
    f.set_variational(np.zeros(f.D), None)
    f.advi_steps(0, 0.1, reset=True)
    for _ in range(3):
        f.advi_steps(100, 0.1)
    ms = sum(f.advi_steps(100, 0.1) for _ in range(5))
    ms_f, ms_w = f.time_lik_kernel(20, True, True), f.time_lik_kernel(50, False, True)
    try:
        ms_c = f.time_lik_kernel(40, 2, True)    # back to back over four rotating copies of Y (input larger than L2), as the main line's roofline
    except Exception:
        ms_c = ms_f
    bl = 4 * M_SPOTS * G_GENES + 4 * M_SPOTS * K_BASIS + 8 * K_BASIS * G_GENES
    t0 = time.perf_counter()
    f.close()
    f = bvg.Fit(phi3, Y3, "nb", "fast", args.engine, **cfg)
    f.run()
    s = f.stats()
    out["cfg3_nb"] = {"workload": "cfg3: 4,992 spots x 3,000 genes, k=20 exp_quad, negative-binomial counts, meanfield ADVI",
                      "value": 500 / (ms * 1e-3), "unit": "evals/s", "us_per_eval": 2.0 * ms,
                      "likelihood_kernel_us_cold_l2": 1e3 * ms_c, "likelihood_kernel_us_flush_before_each": 1e3 * ms_f, "likelihood_kernel_us_l2_resident": 1e3 * ms_w,
                      "roofline_frac_cold": bl / (ms_c * 1e-3) / 1e9 / peak,
                      "fit": {"wall_s": time.perf_counter() - t0, "advi_iters": s["advi_iters"], "mle_iters": s["mle_iters"], "mle_term": s["mle_term"],
                              "mle_fp64_from": s["mle_fp64_from"], "sec_mle": s["sec_mle"], "sec_adapt": s["sec_adapt"], "sec_sga": s["sec_sga"], "converged": s["converged"], "elbo": s["elbo"]}}
    f.close()
    # ---- fp64 engine on cfg2 (north_star's <= 1e-6 path)
    f = bvg.Fit(phi_cfg2, Y_cfg2, "gaussian", "fp64", "simt", **cfg)
    f.set_variational(np.zeros(f.D), None)
    f.advi_steps(0, 0.1, reset=True)
    f.advi_steps(20, 0.1)
    ms = sum(f.advi_steps(50, 0.1) for _ in range(2))
    f.close()
    out["fp64_engine"] = {"workload": "cfg2, precision = fp64 (SIMT engine, fp64 FMA: the tcgen05 path has no fp64 kind)",
                          "value": 100 / (ms * 1e-3), "unit": "evals/s", "us_per_eval": 10.0 * ms}
    # ---- cfg1 stand-in (seu_brain: 2,444 in-tissue spots x 3,000 HVGs): coordinates -> k-means -> basis -> whole default fit
    X1 = synth.scale_coords(synth.masked_blob(synth.visium_grid(), 2444))
    t0 = time.perf_counter()
    C1 = kmeans_centers(X1, K_BASIS, 100, 10, rng=312, device=local_rank)
    phi1 = bvg.basis_build(X1, C1, length_scale_kmeans(C1), "matern", 2.5, device=local_rank)
    t_basis = time.perf_counter() - t0
    Y1, amp1 = synth.expression(phi1, G_GENES, "gaussian", 312)
    runs1 = []
    for _ in range(3):   # median of three fits on fresh handles (see the cfg2 fit block)
        t0 = time.perf_counter()
        f = bvg.Fit(phi1, Y1, "gaussian", "fast", args.engine, **cfg)
        tab = f.run()
        runs1.append((time.perf_counter() - t0 + t_basis, f.stats(), tab))
        f.close()
    runs1.sort(key=lambda r: r[0])
    wall, s, tab = runs1[1]
    blk = {"workload": "stand-in for cfg1 (seu_brain): 2,444 spots x 3,000 genes, k=20 Matern-2.5, gaussian; k-means + basis build + L-BFGS + ADVI + draws",
           "gpu": {"wall_s": wall, "wall_s_runs": [r[0] for r in runs1], "sec_kmeans_basis": t_basis, "sec_h2d": s["sec_h2d"], "sec_mle": s["sec_mle"], "sec_adapt": s["sec_adapt"], "sec_sga": s["sec_sga"],
                   "sec_draws": s["sec_draws"], "advi_iters": s["advi_iters"], "mle_iters": s["mle_iters"], "elbo": s["elbo"], "converged": s["converged"]}}
    if not args.skip_cpu:
        # bounded CPU leg: the SAME pipeline of the fp64 C restatement (all host threads) on the first `gs` genes
        from scipy.stats import spearmanr

        from oracle import oracle as orc
        gs = args.cpu_fit_genes
        nthreads = os.cpu_count() or 1
        t0 = time.perf_counter()
        o = orc.Oracle(phi1, np.asfortranarray(Y1[:, :gs]), "gaussian", nthreads=nthreads)
        th, _ = o.lbfgs(max_iter=1000)
        mu, om, info = o.advi(th, iter=3000, elbo_samples=150, seed=312)
        ref = o.summary(mu, om, 1000, 312)
        cw = time.perf_counter() - t0
        f = bvg.Fit(phi1, np.asfortranarray(Y1[:, :gs]), "gaussian", "fast", args.engine, **cfg)
        t0 = time.perf_counter()
        tg = f.run()
        gw = time.perf_counter() - t0
        f.close()
        top = gs // 3
        blk["cpu_port_sample"] = {"genes": gs, "wall_s": cw, "threads": nthreads, "advi_iters": info["iters"], "eta": info["eta"],
                                  "gpu_wall_s_same_sample": gw,
                                  "spearman_amplitude_mean": float(spearmanr(tg[:, 0], ref[:, 0]).correlation),
                                  "top_third_overlap": len(set(np.argsort(-tg[:, 0])[:top]) & set(np.argsort(-ref[:, 0])[:top])) / top,
                                  "full_problem_extrapolated_wall_s": cw * G_GENES / gs,
                                  "note": "the CPU port's cost per iteration is linear in the gene count; the full 3,000-gene CPU fit was measured once at "
                                          "68.9 s on 16 threads (profiles/r1_fit_agreement_cfg1.json)"}
        blk["speedup_vs_cpu_port_extrapolated"] = blk["cpu_port_sample"]["full_problem_extrapolated_wall_s"] / wall
    blk["spearman_vs_planted_amplitude"] = float(__import__("scipy.stats", fromlist=["spearmanr"]).spearmanr(tab[:, 0], amp1).correlation)
    out["cfg1_fit"] = blk
    return out


def run_ours(args, rank, world, local_rank):
    import torch

    import bayesvg_b200 as bvg
    from bayesvg_b200 import dist as bdist

    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as td
        if os.environ.get("NCCL_DEBUG", "VERSION") == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner out of stdout: rank 0 prints ONE JSON line
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    parity = None
    if world > 1 and not args.skip_parity:
        parity = parity_check(rank, world, local_rank, args.engine, args.comm)
        if not parity["ok"]:
            raise SystemExit(f"parity check failed: {parity}")
    cfg4 = args.workload == "cfg4"
    if cfg4:   # strong scaling: this rank's contiguous share of the 10,000 genes
        g_lo, g_hi = bdist.Shard(rank, world).gene_range(CFG4_G)
        G_local, G_total, g_off, nu, M_w = g_hi - g_lo, CFG4_G, g_lo, 1.5, CFG4_M
    else:      # weak scaling: 3,000 genes on every rank
        G_local, G_total, g_off, nu, M_w = G_GENES, G_GENES * world, G_GENES * rank, 2.5, M_SPOTS
    phi, Y = make_dataset(lambda X, C, ls: bvg.basis_build(X, C, ls, "matern", nu, device=local_rank), 312 + rank,
                          args.workload, G_local)
    E = args.evals_per_step
    cfg = dict(device=local_rank, seed=312)
    if world > 1:
        fit = bdist.make_sharded_fit(phi, Y, bdist.Shard(rank, world), G_total, g_off, "gaussian",
                                     "fast", args.engine, cfg, args.comm)
    else:
        fit = bvg.Fit(phi, Y, "gaussian", "fast", args.engine, **cfg)
    D = fit.D
    fit.set_variational(np.zeros(D), None)
    fit.advi_steps(0, 0.1, reset=True)

    def barrier():
        if world > 1:
            import torch.distributed as td
            td.barrier()
        torch.cuda.synchronize()

    # ---- device-timed steps ----
    for _ in range(max(args.warmup, 3)):
        fit.advi_steps(E, 0.1)
    sampler = ClockSampler(local_rank)
    launches0 = fit.stats()["kernel_launches"]
    barrier()
    sampler.start()
    ms_total = 0.0
    for _ in range(args.steps):
        if not args.no_flush:
            fit.flush_l2()
            if world > 1:
                barrier()   # the flush is a measurement device, not work: without this the first exchange of a step also times the other ranks' flushes
        ms_total += fit.advi_steps(E, 0.1)  # CUDA events on the engine's stream
    barrier()
    sampler.stop_flag = True
    sampler.join()
    launches = fit.stats()["kernel_launches"] - launches0
    if world > 1:
        import torch.distributed as td
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_total = float(t.item())
    # cfg2 (weak): every rank evaluates its own 3,000 genes -> world 3,000-gene evaluations per global evaluation;
    # cfg4 (strong): one evaluation = all 10,000 genes, whatever the number of ranks
    value = (1 if cfg4 else world) * args.steps * E / (ms_total * 1e-3)

    # ---- the dominant kernel against its roofline ----
    peak, peak_src = load_peaks()
    KP = 24
    bytes_launch = 4 * M_w * G_local + 4 * M_w * K_BASIS + 8 * K_BASIS * G_local
    ms_flush = fit.time_lik_kernel(20, True, True)      # L2 flushed before EVERY launch, an event pair around each single launch
    ms_warm = fit.time_lik_kernel(50, False, True)
    # the roofline number: launches back to back over FOUR rotating copies of Y (240 MB at cfg2 > 126 MB of L2; launch i reads
    # copy i mod 4), one event pair around all of them -- every launch streams its Y from HBM, and what is timed is the
    # kernel as it runs inside an evaluation chain, not the ~4 us of launch latency an event pair around ONE 15 us kernel adds
    ms_cold = ms_flush
    cold_how = "L2 flushed (512 MB written + read) before every launch, event pair around each launch"
    if st_engine(fit) == "tcgen05":
        try:
            ms_cold = fit.time_lik_kernel(40, 2, True)
            cold_how = "40 launches back to back over 4 rotating copies of Y (input 4 x Y > L2, no flush inside the timed region), one event pair"
        except Exception:
            pass
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    st = fit.stats()
    eng_name = {1: "simt", 2: "tcgen05"}.get(st["engine_used"], "?")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(eng_name)
    roofline = {"bound": "hbm", "kernel": f"k_lik_{eng_name} (forward+backward, cold L2)",
                "achieved": bytes_launch / (ms_cold * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": bytes_launch / (ms_cold * 1e-3) / 1e9 / peak, "traffic": traffic,
                "traffic_source": (json.load(open(tpath)).get("source") if os.path.exists(tpath) else None), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_launch, "ms_per_launch_cold_l2": ms_cold, "cold_l2_method": cold_how,
                "ms_per_launch_flush_before_each": ms_flush, "frac_flush_before_each": bytes_launch / (ms_flush * 1e-3) / 1e9 / peak,
                "ms_per_launch_l2_resident": ms_warm,
                "l2_resident_equiv_gbs": bytes_launch / (ms_warm * 1e-3) / 1e9}

    # ---- end to end through the C ABI with host buffers (pinned), rank-local ----
    # Y.T is C-contiguous [G][M]: pin that and view it back as the column-major M x G matrix R would pass
    Yh = torch.from_numpy(np.ascontiguousarray(Y.T)).pin_memory().numpy().T if args.pin else Y
    assert Yh.flags.f_contiguous
    e2e_steps = max(2, min(args.steps, 5))
    if world > 1:   # the NCCL communicator belongs to the fit object and is created once, like a user would
        f2 = bdist.make_sharded_fit(phi, Yh, bdist.Shard(rank, world), G_total, g_off,
                                    "gaussian", "fast", args.engine, cfg, args.comm)
    else:
        f2 = bvg.Fit(phi, Yh, "gaussian", "fast", args.engine, **cfg)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        f2.set_y(Yh)                                                       # H2D of this step's Y (+ staging to fp32)
        f2.set_variational(np.zeros(D), None)
        f2.advi_steps(0, 0.1, reset=True)
        f2.advi_steps(E, 0.1)
        mu, om = f2.get_variational()                                      # D2H of the step's result
    barrier()
    e2e_dt = time.perf_counter() - t0
    f2.close()
    if world > 1:
        import torch.distributed as td
        t = torch.tensor([e2e_dt], device="cuda", dtype=torch.float64)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e = {"value": (1 if cfg4 else world) * e2e_steps * E / e2e_dt, "unit": "evals/s",
           "h2d_bytes_per_step": int(Y.nbytes + 8 * D), "d2h_bytes_per_step": int(16 * D),
           "what": f"bvg_fit_set_y_f64(host Y) + {E} ELBO-gradient evals + read-back of (mu, omega); the H2D of Y is "
                   f"amortised over {E} evals here, over >= 3,250 in a real fit"}

    # ---- full default fit (the other half of the metric: SVG fit wall-clock), N = 1 only ----
    fit_info = None
    if cfg4 and not args.skip_fit:   # the whole default fit of the 50k x 10k problem, gene-sharded over the ranks
        barrier()
        t0 = time.perf_counter()
        f3 = (bdist.make_sharded_fit(phi, Y, bdist.Shard(rank, world), G_total, g_off, "gaussian", "fast", args.engine, cfg, args.comm)
              if world > 1 else bvg.Fit(phi, Y, "gaussian", "fast", args.engine, **cfg))
        f3.run()
        barrier()
        s3 = f3.stats()
        fit_info = {"wall_s": time.perf_counter() - t0, "grad_evals": s3["grad_evals"], "elbo_evals": s3["elbo_evals"],
                    "advi_iters": s3["advi_iters"], "eta": s3["eta"], "mle_iters": s3["mle_iters"], "mle_fevals": s3["mle_fevals"],
                    "sec_mle": s3["sec_mle"], "sec_adapt": s3["sec_adapt"], "sec_sga": s3["sec_sga"], "sec_draws": s3["sec_draws"],
                    "converged": s3["converged"], "elbo": s3["elbo"]}
        f3.close()
    if world == 1 and not args.skip_fit and not cfg4:
        # the fit's wall-clock starts at the scaled coordinates: k-means centroids (10 starts) and the basis build
        # (kernel matrix + pivoted QR) on the device, H2D + staging of Y, L-BFGS, eta adaptation, SGA, draws + summaries
        from bayesvg_b200 import synth
        from bayesvg_b200.api import kmeans_centers, length_scale_kmeans
        Xf = synth.scale_coords(synth.visium_grid())
        t_kb_calls = []
        for _ in range(2):   # the first call in a process also loads the k-means / QR kernels (lazy module loading, ~14 ms); the fits are timed warm too
            t0 = time.perf_counter()
            Cf = kmeans_centers(Xf, K_BASIS, 100, 10, rng=312, device=local_rank)
            phif = bvg.basis_build(Xf, Cf, length_scale_kmeans(Cf), "matern", 2.5, device=local_rank)
            t_kb_calls.append(time.perf_counter() - t0)
        t_kb = t_kb_calls[-1]
        Yf, _ = synth.expression(phif, G_GENES, "gaussian", 312)
        # five whole fits, each on a fresh handle (allocation, H2D, staging, L-BFGS, ADVI, draws); the line quotes the median and
        # lists all five: single runs on these boxes carry host-side outliers of 0.1 - 0.3 s (a slow first H2D after an idle
        # period, a stalled driver call) that say nothing about the path
        runs = []
        for _ in range(5):
            t0 = time.perf_counter()
            f3 = bvg.Fit(phif, Yf, "gaussian", "fast", args.engine, **cfg)
            f3.run()
            w3 = time.perf_counter() - t0 + t_kb
            st3 = f3.stats()
            tc0 = time.perf_counter()
            f3.close()
            st3["sec_close"] = time.perf_counter() - tc0   # releasing the handle (cudaFree x 25): 3 ms typically, up to 0.7 s when the driver trims
            runs.append((w3, st3))
        runs.sort(key=lambda r: r[0])
        w3, s3 = runs[len(runs) // 2]
        fit_info = {"wall_s": w3, "wall_s_runs": [r[0] for r in runs],
                    "runs_detail": [{q: round(r[1][q], 4) for q in ("sec_h2d", "sec_mle", "sec_adapt", "sec_sga", "sec_draws", "sec_total", "sec_close")} for r in runs],
                    "sec_kmeans_basis": t_kb, "sec_kmeans_basis_calls": t_kb_calls, "sec_h2d": s3["sec_h2d"], "grad_evals": s3["grad_evals"],
                    "elbo_evals": s3["elbo_evals"], "advi_iters": s3["advi_iters"], "eta": s3["eta"], "mle_iters": s3["mle_iters"],
                    "mle_fevals": s3["mle_fevals"], "mle_term": s3["mle_term"], "sec_mle": s3["sec_mle"], "sec_adapt": s3["sec_adapt"],
                    "sec_sga": s3["sec_sga"], "sec_draws": s3["sec_draws"], "converged": s3["converged"], "elbo": s3["elbo"],
                    "includes": "k-means (10 starts) + basis build + H2D + L-BFGS + eta adaptation + SGA + 1,000 draws + summaries; median of five fits on fresh handles"}
        # the library's default takes the ELBO estimates of a gaussian fit from per-gene sufficient statistics (the same draws and
        # values, O(k^2 G) per draw); the same fit with the DENSE batched ELBO kernel every other model uses, for comparison
        t0 = time.perf_counter()
        f4 = bvg.Fit(phif, Yf, "gaussian", "fast", args.engine, elbo_suffstat=0, **cfg)
        f4.run()
        s4 = f4.stats()
        fit_info["elbo_path"] = "sufficient statistics (library default for gaussian, k <= 30)" if cfg.get("elbo_suffstat", bvg.engine.default_config().elbo_suffstat) else "dense"
        fit_info["dense_elbo"] = {"wall_s": time.perf_counter() - t0 + t_kb, "sec_mle": s4["sec_mle"], "sec_adapt": s4["sec_adapt"],
                                  "sec_sga": s4["sec_sga"], "advi_iters": s4["advi_iters"], "eta": s4["eta"],
                                  "elbo": s4["elbo"], "elbo_default_path": s3["elbo"]}
        f4.close()

    # ---- the same kernel on an input that does NOT fit in L2: one config-4 shard (50,000 spots x 1,250 genes =
    # what each of 8 GPUs holds at 50k x 10k; Y = 250 MB), back-to-back launches, no flush needed ----
    roofline_large = None
    if world == 1 and not args.skip_large and not cfg4:
        from bayesvg_b200 import synth
        Ml, Gl = 50000, 1250
        Xl = synth.scale_coords(synth.random_points(Ml, 1))
        Cl = synth.det_kmeans(Xl, K_BASIS, 312, iters=30)
        from bayesvg_b200.api import length_scale_kmeans
        phil = bvg.basis_build(Xl, Cl, length_scale_kmeans(Cl), "matern", 1.5, device=local_rank)
        Yl, _ = synth.expression(phil, Gl, "gaussian", 312)
        fl = bvg.Fit(phil, Yl, "gaussian", "fast", args.engine, **cfg)
        fl.set_variational(np.zeros(fl.D), None)
        fl.advi_steps(5, 0.1, reset=True)
        ms_l = fl.time_lik_kernel(20, False, True)
        ms_eval_l = fl.advi_steps(50, 0.1) / 50
        bl = 4 * Ml * Gl + 4 * Ml * K_BASIS + 8 * K_BASIS * Gl
        roofline_large = {"workload": "one cfg4 shard: 50,000 spots x 1,250 genes (Y = 250 MB > L2), k=20 Matern-1.5",
                          "bound": "hbm", "achieved": bl / (ms_l * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": bl / (ms_l * 1e-3) / 1e9 / peak, "ms_per_launch": ms_l,
                          "algorithmic_bytes_per_launch": bl, "ms_per_elbo_grad_eval": ms_eval_l}
        fl.close()

    # ---- the large-basis kernels (BASELINE config 5: k ~ 400) against the TENSOR roofline: two tcgen05 GEMM kernels,
    # BF16x3 (every product is 3 bf16 MMAs), 50,000 spots x 1,280 genes = a quarter of a config-5 shard ----
    roofline_kbig = None
    if world == 1 and not args.skip_large and not cfg4:
        from bayesvg_b200 import synth
        Mb, Gb, kb = 50000, 1280, 400
        Xb = synth.scale_coords(synth.random_points(Mb, 2))
        Cb = bvg.kmeans_centers(Xb, kb, 30, 1, rng=312, device=local_rank)
        from bayesvg_b200.api import length_scale_kmeans
        phib = bvg.basis_build(Xb, Cb, length_scale_kmeans(Cb), "matern", 1.5, device=local_rank)
        rngb = np.random.default_rng(312)
        Yb = np.asfortranarray((phib @ rngb.normal(size=(kb, Gb))) * np.sqrt(Mb / kb) * 0.3 + rngb.normal(size=(Mb, Gb)))
        Yb = (Yb - Yb.mean()) / Yb.std(ddof=1)
        fb = bvg.Fit(phib, Yb, "gaussian", "fast", args.engine, **cfg)
        fb.set_variational(np.zeros(fb.D), None)
        fb.advi_steps(3, 0.1, reset=True)
        ms_b = fb.time_lik_kernel(20, False, True)
        ms_eval_b = fb.advi_steps(20, 0.1) / 20
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        tpeak = float(peaks.get("bf16_tflops_sustained", 1418.0))
        executed = 3 * 4.0 * (kb + 1) * Mb * Gb / (ms_b * 1e-3) / 1e12
        roofline_kbig = {"workload": "50,000 spots x 1,280 genes, k = 400 (quarter of a config-5 shard), gaussian", "bound": "tensor",
                         "kernel": "k_tcbig<FWD> + k_tcbig<BWD> (BF16x3 tcgen05 GEMMs)", "achieved": executed, "peak": tpeak,
                         "unit": "TFLOP/s", "frac": executed / tpeak, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained",
                         "algorithmic_tflops": (4.0 * kb * Mb * Gb + 12.0 * Mb * Gb) / (ms_b * 1e-3) / 1e12,
                         "ms_per_pass": ms_b, "ms_per_elbo_grad_eval": ms_eval_b,
                         "what": "achieved = executed bf16 MMA flops (3 MMAs per product x 2 GEMMs x 2 flop); algorithmic = 4 k M G + 12 M G per pass"}
        fb.close()

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu and not cfg4:
        nthreads = os.cpu_count() or 1
        n, dt = cpu_evals(phi, Y, args.cpu_seconds, 3, nthreads)
        cpu = {"value": n / dt, "unit": "evals/s", "cores": nthreads, "kind": "port",
               "sample": f"{n} ELBO-gradient evals of the full 4,992 x 3,000 workload in {dt:.1f} s (fp64 C restatement, OpenMP)"}
        # the reference's meanfield fit is single-threaded (stan_threads = FALSE, findSpatiallyVariableFeaturesBayes.R:123-129)
        n1, dt1 = cpu_evals(phi, Y, 3.0, 3, 1)
        cpu["single_thread"] = {"value": n1 / dt1, "unit": "evals/s", "cores": 1,
                                "sample": f"{n1} evals in {dt1:.1f} s on one thread (what stan_threads = FALSE gives the reference)"}
    extra = {}
    if rank == 0 and world == 1 and not cfg4 and not args.skip_extra:
        extra = extra_blocks_n1(args, local_rank, phi, Y)
    fit.close()
    strong = None
    if not cfg4 and not args.skip_strong:
        del Y
        strong = strong_cfg4(args, rank, world, local_rank, barrier)
    if rank == 0:
        print(json.dumps({
            "metric": "elbo_grad_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong" if cfg4 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": ("cfg4: 50,000 spots x 10,000 genes in total (gene-sharded), k=20 Matern-1.5, gaussian, meanfield ADVI"
                                    if cfg4 else "cfg2: 4,992 spots x 3,000 genes per GPU, k=20 Matern-2.5, gaussian, meanfield ADVI"),
                       "evals_per_step": E, "engine": eng_name,
                       "l2": ("L2 flushed (512 MB write) before every timed step; inside a step the evals re-read "
                              "the 60 MB Y as the real fit does" if not args.no_flush else "no flush"),
                       "genes_total": G_total,
                       "exchange": (None if world == 1 else
                                    ("2k+10 fp64 sums per eval, exchanged inside the persistent ADVI kernel over CUDA-IPC peer memory (NVLink): CTA 0 pushes, every CTA polls and sums in rank order"
                                     if args.comm == "peer" else "2k+10 fp64 sums per eval, ncclAllReduce between k_gene and k_finish"))},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "gpu_launches_note": "kernel launches inside the timed region: k_prep + k_advi_persist (one cooperative launch = the whole stretch of evals_per_step evaluations) per step",
            "roofline": roofline,
            "roofline_input_larger_than_l2": roofline_large, "roofline_large_basis": roofline_kbig, "cpu_baseline": cpu, "fit": fit_info,
            "strong_cfg4": strong, "parity_check": parity, **extra,
        }))
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--evals-per-step", type=int, default=100)
    ap.add_argument("--ref-evals-per-step", type=int, default=2)
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"])
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--pin", action="store_true", default=True)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-fit", action="store_true")
    ap.add_argument("--skip-large", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--cpu-fit-genes", type=int, default=192, help="genes of the bounded CPU fit leg of cfg1_fit")
    ap.add_argument("--skip-strong", action="store_true", help="no strong_cfg4 block")
    ap.add_argument("--skip-extra", action="store_true", help="no cfg3_nb / fp64_engine / cfg1_fit blocks (N = 1)")
    ap.add_argument("--skip-parity", action="store_true", help="no sharded-vs-unsharded pre-flight (N > 1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

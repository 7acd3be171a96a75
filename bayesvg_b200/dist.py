"""Gene sharding across GPUs: one process per GPU (torchrun), contiguous gene ranges, one
all-reduce of 2k+10 doubles per evaluation inside the library (csrc/comm.cu): by default fused into
the per-gene kernel over CUDA-IPC peer memory (NVLink), optionally NCCL between kernels.
torch.distributed is used only as plumbing: to exchange the IPC handles / the NCCL unique id and to
gather the per-shard summary tables."""
from dataclasses import dataclass

import numpy as np


@dataclass
class Shard:
    rank: int
    world_size: int

    def gene_range(self, G):
        """contiguous split; the first G % world ranks hold one extra gene"""
        base, rem = divmod(G, self.world_size)
        start = self.rank * base + min(self.rank, rem)
        return start, start + base + (1 if self.rank < rem else 0)


def gene_ranges(G, world_size):
    return [Shard(r, world_size).gene_range(G) for r in range(world_size)]


def _bcast_bytes(b, src=0):
    import torch
    import torch.distributed as td
    dev = "cuda" if td.get_backend() == "nccl" else "cpu"
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if td.get_rank() == src:
        t[: len(b)] = torch.tensor(list(b), dtype=torch.uint8)
    td.broadcast(t, src)
    return bytes(t.cpu().tolist())


def _allgather_bytes(b):
    """every rank contributes len(b) bytes; returns the concatenation in rank order"""
    import torch
    import torch.distributed as td
    dev = "cuda" if td.get_backend() == "nccl" else "cpu"
    mine = torch.tensor(list(b), dtype=torch.uint8, device=dev)
    out = [torch.zeros_like(mine) for _ in range(td.get_world_size())]
    td.all_gather(out, mine)
    return b"".join(bytes(t.cpu().tolist()) for t in out)


def make_sharded_fit(phi, Y_local, shard, G_total, gene_offset, likelihood, precision, engine, cfg, comm="peer",
                     gene_depths_local=None):
    """Create this rank's Fit and connect it to the other shards (collective: every rank must call).
    comm = "peer": in-kernel all-reduce over CUDA-IPC peer memory; "nccl": ncclAllReduce between kernels."""
    from . import engine as E
    if comm not in ("peer", "nccl"):
        raise ValueError("comm must be 'peer' or 'nccl'")
    fit = E.Fit(phi, Y_local, likelihood, precision, engine, gene_depths=gene_depths_local, rank=shard.rank,
                world_size=shard.world_size, gene_offset=gene_offset, G_total=G_total, **cfg)
    if shard.world_size == 1:
        return fit
    if comm == "peer":
        fit.peer_init(_allgather_bytes(fit.peer_handle()))
    else:
        uid = E.comm_unique_id() if shard.rank == 0 else b""
        fit.comm_init(_bcast_bytes(uid)[:128])
    # every rank has mapped every peer's exchange buffer before anyone's first exchange can start
    import torch.distributed as td
    td.barrier()
    return fit


def fit_sharded(phi, Y, shard, likelihood, precision, engine, cfg, comm="peer", gene_depths=None, save_model=False,
                n_draws=None, genes=None, extra=None):
    """Y: full M x G on every rank (each keeps its slice).  Returns (gathered G x 8 table, model).
    save_model (save.model = TRUE, R/findSpatiallyVariableFeaturesBayes.R:483-489): every rank draws its columns of
    the posterior matrix (bvg_fit_get_posterior on a sharded fit), the columns are gathered and put into the
    declaration order of the UNSHARDED model, and every rank returns the same api.VariationalFit."""
    import torch
    import torch.distributed as td
    G = Y.shape[1]
    g0, g1 = shard.gene_range(G)
    fit = make_sharded_fit(phi, np.asfortranarray(Y[:, g0:g1]), shard, G, g0, likelihood, precision, engine, cfg, comm,
                           None if gene_depths is None else np.asarray(gene_depths)[g0:g1])
    local = fit.run()
    dev = "cuda" if td.get_backend() == "nccl" else "cpu"
    model = None
    if save_model:
        model = _gather_model(fit, shard, likelihood, G, g0, g1, gene_depths is not None, n_draws, genes, extra or {}, dev)
    fit.close()
    sizes = [b - a for a, b in gene_ranges(G, shard.world_size)]
    bufs = [torch.zeros(s, 8, dtype=torch.float64, device=dev) for s in sizes]
    td.all_gather(bufs, torch.tensor(np.ascontiguousarray(local), device=dev)) if len(set(sizes)) == 1 else _gather_ragged(bufs, local, dev, sizes)
    table = np.concatenate([b.cpu().numpy() for b in bufs], 0)
    return table, model


def assemble_posterior(parts, likelihood, G, k, ranges, depth=False):
    """parts[r] = (n x (3 + D_r + ntp G_r)) posterior matrix of the shard holding genes ranges[r] (local declaration
    order, then beta0[G_r] (hierarchical intercept only), amplitude_sq[G_r]) -> the n x (3 + D + ntp G) matrix of the
    unsharded model.  lp__/log_p__/log_g__ and the replicated gene-shared columns are identical on every shard."""
    LG = layout(likelihood, G, k, depth)
    D, ntp = LG["D"], (1 if depth else 2)
    n = parts[0].shape[0]
    out = np.empty((n, 3 + D + ntp * G), order="F")
    out[:, :3] = parts[0][:, :3]
    for (g0, g1), P in zip(ranges, parts):
        Gl = g1 - g0
        Dl = layout(likelihood, Gl, k, depth)["D"]
        if P.shape != (n, 3 + Dl + ntp * Gl):
            raise ValueError("shard matrix has the wrong shape")
        out[:, 3 + shard_index(likelihood, G, k, g0, g1, depth)] = P[:, 3:3 + Dl]
        for t in range(ntp):   # beta0[G] (if any), then amplitude_sq[G]
            out[:, 3 + D + t * G + g0:3 + D + t * G + g1] = P[:, 3 + Dl + t * Gl:3 + Dl + (t + 1) * Gl]
    return out


def _gather_model(fit, shard, likelihood, G, g0, g1, depth, n_draws, genes, extra, dev):
    import torch
    import torch.distributed as td
    from . import engine as E
    from .api import VariationalFit
    n = int(fit.cfg.n_draws if n_draws is None else n_draws)
    k = fit.k
    local = np.ascontiguousarray(fit.posterior(n))
    mu, om = fit.get_variational()
    ranges = gene_ranges(G, shard.world_size)
    widths = [3 + layout(likelihood, b - a, k, depth)["D"] + (1 if depth else 2) * (b - a) for a, b in ranges]
    wmax = max(widths)

    def gather(mat):   # ragged in the column count: pad to the widest shard
        pad = torch.zeros(mat.shape[0], wmax, dtype=torch.float64, device=dev)
        pad[:, :mat.shape[1]] = torch.tensor(mat, device=dev)
        outs = [torch.zeros_like(pad) for _ in ranges]
        td.all_gather(outs, pad)
        return [o.cpu().numpy() for o in outs]

    parts = [p[:, :w] for p, w in zip(gather(local), widths)]
    draws = assemble_posterior(parts, likelihood, G, k, ranges, depth)
    mo = gather(np.stack([np.concatenate([np.zeros(3), mu]), np.concatenate([np.zeros(3), om])]))
    LGD = layout(likelihood, G, k, depth)["D"]
    mu_g, om_g = np.empty(LGD), np.empty(LGD)
    for (a, b), m2 in zip(ranges, mo):
        Dl = layout(likelihood, b - a, k, depth)["D"]
        idx = shard_index(likelihood, G, k, a, b, depth)
        mu_g[idx], om_g[idx] = m2[0, 3:3 + Dl], m2[1, 3:3 + Dl]
    # column names of the unsharded model: ask a name-only helper with the global G
    cols = E.posterior_column_names(likelihood == "nb", depth, G, k)
    return VariationalFit(cols, draws, mu=mu_g, omega=om_g, trace=fit.trace(), stats=fit.stats(), genes=genes, **extra)


def _gather_ragged(bufs, local, dev, sizes):
    import torch
    import torch.distributed as td
    mx = max(sizes)
    pad = torch.zeros(mx, 8, dtype=torch.float64, device=dev)
    pad[: local.shape[0]] = torch.tensor(np.ascontiguousarray(local), device=dev)
    out = [torch.zeros(mx, 8, dtype=torch.float64, device=dev) for _ in sizes]
    td.all_gather(out, pad)
    for b, o, s in zip(bufs, out, sizes):
        b.copy_(o[:s])


def layout(likelihood, G, k, depth=False):
    """slot offsets of the unconstrained vector (Stan declaration order; inst/approxGP.stan:12-23,
    inst/approxGP4.stan:12-23 of the reference; depth=True: inst/approxGP2.stan:13-23 / approxGP3.stan:13-23, where
    mu_beta0 / sigma_beta0 name the slots of beta0 / beta1 and there is no z_beta0)"""
    if depth:
        return dict(mu_beta0=0, sigma_beta0=1, tail=2, amplitude=3, mu_amplitude=3 + G, sigma_amplitude=4 + G,
                    mu_alpha=5 + G, sigma_alpha=5 + G + k, z_alpha_t=5 + G + 2 * k, D=5 + G + 2 * k + k * G)
    if likelihood == "gaussian":
        o = dict(mu_beta0=0, sigma_beta0=1, z_beta0=2, mu_amplitude=2 + G, sigma_amplitude=3 + G, amplitude=4 + G,
                 mu_alpha=4 + 2 * G, sigma_alpha=4 + 2 * G + k, z_alpha_t=4 + 2 * G + 2 * k, tail=4 + 2 * G + 2 * k + k * G)
    else:
        o = dict(mu_beta0=0, sigma_beta0=1, z_beta0=2, tail=2 + G, amplitude=3 + G, mu_amplitude=3 + 2 * G,
                 sigma_amplitude=4 + 2 * G, mu_alpha=5 + 2 * G, sigma_alpha=5 + 2 * G + k, z_alpha_t=5 + 2 * G + 2 * k)
    o["D"] = 5 + 2 * G + 2 * k + k * G
    return o


def shard_index(likelihood, G_total, k, g0, g1, depth=False):
    """indices into the GLOBAL unconstrained vector that make up the LOCAL vector of the shard holding genes
    [g0, g1): local = global[shard_index(...)].  The 2k+5 gene-shared entries appear in every shard."""
    Gl = g1 - g0
    LG, LL = layout(likelihood, G_total, k, depth), layout(likelihood, Gl, k, depth)
    idx = np.empty(LL["D"], dtype=np.int64)
    for name in ("mu_beta0", "sigma_beta0", "mu_amplitude", "sigma_amplitude", "tail"):
        idx[LL[name]] = LG[name]
    idx[LL["mu_alpha"]:LL["mu_alpha"] + k] = LG["mu_alpha"] + np.arange(k)
    idx[LL["sigma_alpha"]:LL["sigma_alpha"] + k] = LG["sigma_alpha"] + np.arange(k)
    if not depth:
        idx[LL["z_beta0"]:LL["z_beta0"] + Gl] = LG["z_beta0"] + np.arange(g0, g1)
    idx[LL["amplitude"]:LL["amplitude"] + Gl] = LG["amplitude"] + np.arange(g0, g1)
    idx[LL["z_alpha_t"]:LL["z_alpha_t"] + k * Gl] = LG["z_alpha_t"] + np.arange(k * g0, k * g1)
    return idx


def shared_mask(likelihood, G, k, depth=False):
    """boolean mask of the 2k+5 gene-shared entries of a (local or global) vector"""
    L = layout(likelihood, G, k, depth)
    m = np.zeros(L["D"], dtype=bool)
    for name in ("mu_beta0", "sigma_beta0", "mu_amplitude", "sigma_amplitude", "tail"):
        m[L[name]] = True
    m[L["mu_alpha"]:L["mu_alpha"] + 2 * k] = True
    return m

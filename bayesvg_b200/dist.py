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
    return fit


def fit_sharded(phi, Y, shard, likelihood, precision, engine, cfg, comm="peer", gene_depths=None):
    """Y: full M x G on every rank (each keeps its slice).  Returns the gathered G x 8 table."""
    import torch
    import torch.distributed as td
    G = Y.shape[1]
    g0, g1 = shard.gene_range(G)
    fit = make_sharded_fit(phi, np.asfortranarray(Y[:, g0:g1]), shard, G, g0, likelihood, precision, engine, cfg, comm,
                           None if gene_depths is None else np.asarray(gene_depths)[g0:g1])
    local = fit.run()
    fit.close()
    dev = "cuda" if td.get_backend() == "nccl" else "cpu"
    sizes = [b - a for a, b in gene_ranges(G, shard.world_size)]
    bufs = [torch.zeros(s, 8, dtype=torch.float64, device=dev) for s in sizes]
    td.all_gather(bufs, torch.tensor(np.ascontiguousarray(local), device=dev)) if len(set(sizes)) == 1 else _gather_ragged(bufs, local, dev, sizes)
    table = np.concatenate([b.cpu().numpy() for b in bufs], 0)
    return table, None


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

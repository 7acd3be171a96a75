"""Gene sharding across GPUs: one process per GPU (torchrun), contiguous gene ranges, one
all-reduce of 2k+10 doubles per evaluation inside the library (NCCL, csrc/comm.cu).
torch.distributed is used only as plumbing: to broadcast the NCCL unique id and gather the
per-shard summary tables."""
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


def make_sharded_fit(phi, Y_local, shard, G_total, gene_offset, likelihood, precision, engine, cfg):
    """Create this rank's Fit and join the NCCL communicator (collective: every rank must call)."""
    from . import engine as E
    uid = E.comm_unique_id() if shard.rank == 0 else b""
    uid = _bcast_bytes(uid)[:128]
    fit = E.Fit(phi, Y_local, likelihood, precision, engine, rank=shard.rank, world_size=shard.world_size,
                gene_offset=gene_offset, G_total=G_total, **cfg)
    fit.comm_init(uid)
    return fit


def fit_sharded(phi, Y, shard, likelihood, precision, engine, cfg):
    """Y: full M x G on every rank (each keeps its slice).  Returns the gathered G x 8 table."""
    import torch
    import torch.distributed as td
    G = Y.shape[1]
    g0, g1 = shard.gene_range(G)
    fit = make_sharded_fit(phi, np.asfortranarray(Y[:, g0:g1]), shard, G, g0, likelihood, precision, engine, cfg)
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

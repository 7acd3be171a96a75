"""CPU, world_size 2, gloo: the decomposition the multi-GPU path relies on.  Each rank evaluates the
log density on ITS genes only (here with the oracle standing in for the GPU shard); summing the
per-shard results with one all-reduce and removing the (P-1) extra copies of the gene-shared prior
terms must reproduce the unsharded log density and its gradient.  Also exercises the torch.distributed
plumbing of bayesvg_b200.dist (byte broadcast, ragged gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from bayesvg_b200 import dist
from oracle import oracle as orc
from tests.util import make_problem


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shared_prior(theta, lik, G, k):
    """log prior + jacobian of the 2k+5 gene-shared parameters (propto form) and its gradient"""
    L = dist.layout(lik, G, k)
    g = np.zeros_like(theta)
    mb, lsb, ma, lsa, lt = (theta[L[n]] for n in ("mu_beta0", "sigma_beta0", "mu_amplitude", "sigma_amplitude", "tail"))
    lp = -mb ** 2 / 8 - np.exp(2 * lsb) / 2 - ma ** 2 / 8 - np.exp(2 * lsa) / 2 + lsb + lsa + lt
    g[L["mu_beta0"]], g[L["sigma_beta0"]] = -mb / 4, -np.exp(2 * lsb) + 1
    g[L["mu_amplitude"]], g[L["sigma_amplitude"]] = -ma / 4, -np.exp(2 * lsa) + 1
    if lik == "gaussian":
        lp += -np.exp(2 * lt) / 2
        g[L["tail"]] = -np.exp(2 * lt) + 1
    else:
        ph = np.exp(lt)
        lp += -np.log1p(ph * ph / 4)
        g[L["tail"]] = -0.5 * ph * ph / (1 + ph * ph / 4) + 1
    a = theta[L["mu_alpha"]:L["mu_alpha"] + k]
    ls = theta[L["sigma_alpha"]:L["sigma_alpha"] + k]
    lp += (-a ** 2 / 8 - np.exp(2 * ls) / 2 + ls).sum()
    g[L["mu_alpha"]:L["mu_alpha"] + k] = -a / 4
    g[L["sigma_alpha"]:L["sigma_alpha"] + k] = -np.exp(2 * ls) + 1
    return lp, g


def _worker(rank, world, port, lik, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M, G, k = 90, 11, 4
        p = make_problem(M, G, k, lik, seed=5)
        theta = np.random.default_rng(9).normal(0, 0.4, dist.layout(lik, G, k)["D"])
        # plumbing: byte broadcast (the NCCL unique id travels this way)
        blob = bytes(range(128)) if rank == 0 else b""
        assert dist._bcast_bytes(blob) == bytes(range(128))
        # ... and the CUDA-IPC handles of the peer-memory exchange are all-gathered in rank order
        assert dist._allgather_bytes(bytes([rank] * 64)) == bytes([0] * 64 + [1] * 64)
        g0, g1 = dist.Shard(rank, world).gene_range(G)
        idx = dist.shard_index(lik, G, k, g0, g1)
        o = orc.Oracle(p["phi"], p["y"][:, g0:g1], lik)
        lp_r, g_r = o.log_prob(theta[idx], True, True)
        sp, sg = _shared_prior(theta[idx], lik, g1 - g0, k)
        # what crosses ranks: the log density and the shared-parameter gradient sums (2k+6 scalars)
        sm = dist.shared_mask(lik, g1 - g0, k)
        buf = torch.tensor(np.concatenate([[lp_r - sp], (g_r - sg)[sm]]))
        td.all_reduce(buf)
        lp = float(buf[0]) + sp
        g_loc = g_r.copy()
        g_loc[sm] = buf[1:].numpy() + sg[sm]
        # ragged gather of per-shard tables
        sizes = [b - a for a, b in dist.gene_ranges(G, world)]
        bufs = [torch.zeros(s, 8, dtype=torch.float64) for s in sizes]
        local = np.full((g1 - g0, 8), float(rank))
        dist._gather_ragged(bufs, local, "cpu", sizes)
        assert all(float(b.mean()) == r for r, b in enumerate(bufs))
        if rank == 0:
            full = orc.Oracle(p["phi"], p["y"], lik)
            rlp, rg = full.log_prob(theta, True, True)
            q.put((lp, rlp, float(np.abs(g_loc - rg[idx]).max()), float(np.abs(rg).max())))
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("lik", ["gaussian", "nb"])
def test_two_rank_gene_sharding_reproduces_the_unsharded_density(lik):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lik, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    lp, rlp, gerr, gmax = q.get(timeout=10)
    assert lp == pytest.approx(rlp, rel=1e-12)
    assert gerr <= 1e-10 * gmax

"""CPU tier, world_size 2 over gloo: the multi-GPU host logic (row sharding keyed by global row index, one all-reduce of
the merged rates / row-block merge of scan maps).  The per-rank compute is played by the oracle on Philox variates
keyed by GLOBAL row ids -- exactly what each GPU rank's kernels generate -- so the test also shows that the result does
not depend on the number of ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alplib_b200 import dist as D


def test_shard_range_partitions():
    for n in (0, 1, 7, 10, 200, 10_000):
        for w in (1, 2, 3, 8):
            blocks = [D.shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    t = np.arange(20).reshape(10, 2)
    rows, off = D.shard_rows(t, 1, 3)
    assert off == 4 and rows.tolist() == t[4:7].tolist()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_rate(pf, row_offset, ma, g, ns, seed):
    from oracle import alp_oracle as O
    from oracle import philox as PX
    keep = O.compton_row_valid(pf, ma)
    us = [PX.uniform_1(seed, PX.TAG_COMPTON_EA, int(r) + row_offset, ns) for r in np.nonzero(keep)[0]]
    if not us:
        return 0.0, np.zeros(0)
    E, F = O.compton_simulate(pf, ma, g, 74, O.AbsCrossSection(O.Material("W")), ns, O.UniformSource(np.concatenate(us)))
    p = O.propagate(E, F, ma, O.W_ee(g, ma), 1.0, 4.0, 0.2, geom_accept=O.geom_acceptance(0.04, 4.0))
    ew, r = O.decays(E, p["decay_w"], 100, 5.0)
    return float(r), E


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        e = np.logspace(np.log10(0.5), 3, 41)
        pf = np.stack([e, 1e16 * e ** -1.5 * np.exp(-e / 150)], axis=1)
        rows, off = D.shard_rows(pf, rank, world)
        r, E = _rank_rate(rows, off, 1.5, 1e-7, 64, 99)
        t = torch.tensor([r, float(E.shape[0])], dtype=torch.float64)
        D.all_reduce_sum(t)
        # scan-map merge: rank r owns a contiguous block of mass rows
        lo, hi = D.shard_range(7, rank, world)
        local = torch.arange(lo * 3, hi * 3, dtype=torch.float64).view(hi - lo, 3)
        full = D.merge_row_blocks(local, lo, 7)
        if rank == 0:
            q.put((t.tolist(), full.tolist(), D.world()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharding_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    (rate2, n2), full, (rk, ws) = q.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    e = np.logspace(np.log10(0.5), 3, 41)
    pf = np.stack([e, 1e16 * e ** -1.5 * np.exp(-e / 150)], axis=1)
    rate1, E1 = _rank_rate(pf, 0, 1.5, 1e-7, 64, 99)
    assert (rk, ws) == (0, 2)
    assert n2 == E1.shape[0]
    assert rate2 == pytest.approx(rate1, rel=1e-13)
    assert full == np.arange(21, dtype=float).reshape(7, 3).tolist()

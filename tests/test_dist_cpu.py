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


def test_shard_kept_rows_balances_the_kept_rows():
    e = np.logspace(np.log10(0.5), 3, 50)
    t = np.stack([e, np.ones(50)], axis=1)
    keep = e >= 4.0
    for w in (1, 2, 3, 8):
        blocks = [D.shard_kept_rows(t, keep, r, w) for r in range(w)]
        kept = [b[b[:, 0] >= 4.0] for b, _ in blocks]
        sizes = [k.shape[0] for k in kept]
        assert sum(sizes) == int(keep.sum()) and max(sizes) - min(sizes) <= 1
        assert np.array_equal(np.concatenate(kept), t[keep])
        for b, off in blocks:
            if b.shape[0]:
                assert np.array_equal(b, t[off:off + b.shape[0]]) and keep[off] and keep[off + b.shape[0] - 1]
    none, off = D.shard_kept_rows(t, np.zeros(50, bool), 1, 4)
    assert none.shape[0] == 0 and off == 0

    class Fl:                                  # shard_flux only touches the table attribute and row_offset
        _TABLE_ATTR = "photon_flux"
        photon_flux = t.tolist()
        row_offset = 3

        def _row_keep(self, table):
            return table[:, 0] >= 4.0
    f = D.shard_flux(Fl(), rank=1, nranks=2)
    b, off = D.shard_kept_rows(t, keep, 1, 2)
    assert np.array_equal(f.photon_flux, b) and f.row_offset == 3 + off


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
        # the same blocks through the all-gather path (ragged: 4 + 3 rows) and the packed [rates || hist] merge
        counts = [b - a for a, b in (D.shard_range(7, r, world) for r in range(world))]
        gathered = torch.cat(D.gather_row_blocks(local, counts))
        rr, hh = D.merged_rates_and_hists([torch.tensor([1.0 + rank], dtype=torch.float64), torch.tensor([10.0 * rank], dtype=torch.float64)],
                                          [torch.arange(5, dtype=torch.float64) * (rank + 1)])
        if rank == 0:
            q.put((t.tolist(), full.tolist(), D.world(), gathered.tolist(), rr.tolist(), hh[0].tolist()))
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
    (rate2, n2), full, (rk, ws), gathered, rr, hh = q.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    e = np.logspace(np.log10(0.5), 3, 41)
    pf = np.stack([e, 1e16 * e ** -1.5 * np.exp(-e / 150)], axis=1)
    rate1, E1 = _rank_rate(pf, 0, 1.5, 1e-7, 64, 99)
    assert (rk, ws) == (0, 2)
    assert n2 == E1.shape[0]
    assert rate2 == pytest.approx(rate1, rel=1e-13)
    assert full == np.arange(21, dtype=float).reshape(7, 3).tolist() and gathered == full
    assert rr == [3.0, 10.0] and hh == [0.0, 3.0, 6.0, 9.0, 12.0]

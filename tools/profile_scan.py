#!/usr/bin/env python
"""Per-stage breakdown of one (m_a, g) scan call (BASELINE config 5 and the 100x larger scan): where the fixed floor of a
small scan goes.  Run single-process or under torchrun (every rank prints nothing but rank 0).
    python tools/profile_scan.py            |  torchrun --nproc-per-node N tools/profile_scan.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A            # noqa: E402
import bench                       # noqa: E402
from alplib_b200 import scan as S  # noqa: E402
from alplib_b200.runtime import Runtime   # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
W = A.Material("W")
ma, g = np.logspace(-2, 2, 200), np.logspace(-9, -3, 200)
kw = dict(days_exposure=100, threshold=5.0, seed=bench.SEED, **bench.GEOM)
out = {"world": world}
for key, n_bins, ns in (("c5", 1000, 100), ("large", 10_000, 1000)):
    pf = bench.c2_spectrum(n_bins)
    my = ma[rank::world]
    res = {}
    for proc, extra in (("compton", dict(n_samples=ns)), ("primakoff", {})):
        for _ in range(3):
            S.scan_rates_distributed(proc, pf, W, ma, g, **extra, **kw)
        torch.cuda.synchronize()
        reps = 10 if key == "c5" else 2
        # (1) the whole distributed call, host clock
        t0 = time.perf_counter()
        for _ in range(reps):
            S.scan_rates_distributed(proc, pf, W, ma, g, **extra, **kw)
        t_call = (time.perf_counter() - t0) / reps
        # (2) this rank's share: launch only (device tensors out), device time by events and host time to queue it
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            d = S.scan_rates(proc, pf, W, my, g, return_device=True, g0=float(g[0]), **extra, **kw)
        e1.record()
        t_queue = (time.perf_counter() - t0) / reps
        torch.cuda.synchronize()
        t_dev = e0.elapsed_time(e1) * 1e-3 / reps
        # (3) device -> host of the full map through pinned memory
        full = torch.empty((200, 200), dtype=torch.float64, device=d.device)
        rt = Runtime.get(d.device)
        t0 = time.perf_counter()
        for _ in range(reps):
            rt.to_host(full)
        t_d2h = (time.perf_counter() - t0) / reps
        entry = {"call_ms": t_call * 1e3, "host_queue_ms": t_queue * 1e3, "device_ms": t_dev * 1e3, "d2h_map_ms": t_d2h * 1e3}
        if world > 1:
            # (4) the gather alone, ranks aligned by a barrier first (launch latency without rank skew): the peer-window kernel
            # the scan uses on one node, and NCCL's all_gather_into_tensor beside it
            from alplib_b200.peer import peer_group
            pg = peer_group(None, d)
            n_loc = (200 + world - 1) // world
            loc = torch.zeros((n_loc, 200), dtype=torch.float64, device=d.device)
            flat = torch.empty((world * n_loc, 200), dtype=torch.float64, device=d.device)
            mine = torch.zeros((len(range(rank, 200, world)), 200), dtype=torch.float64, device=d.device)

            def coll(fn):
                for _ in range(3):
                    fn()
                t_coll = 0.0
                for _ in range(reps):
                    dist.barrier(); torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    fn()
                    torch.cuda.synchronize()
                    t_coll += time.perf_counter() - t0
                return t_coll / reps * 1e3
            entry["nccl_all_gather_ms"] = coll(lambda: dist.all_gather_into_tensor(flat, loc))
            entry["all_gather_ms"] = entry["nccl_all_gather_ms"]
            if pg is not None:
                entry["peer_all_gather_ms"] = coll(lambda: pg.all_gather_rows(mine, rank, world, 200))
                entry["all_gather_ms"] = entry["peer_all_gather_ms"]
            # (5) device time of every rank's share: the slowest rank sets the pace of the collective
            tt = torch.tensor([t_dev * 1e3], dtype=torch.float64, device=d.device)
            allt = torch.empty(world, dtype=torch.float64, device=d.device)
            dist.all_gather_into_tensor(allt, tt)
            entry["device_ms_per_rank"] = [round(float(x), 4) for x in allt.cpu()]
        entry["unaccounted_ms"] = (t_call - max(t_queue, max(entry.get("device_ms_per_rank", [t_dev * 1e3])) * 1e-3) - t_d2h
                                   - entry.get("all_gather_ms", 0.0) * 1e-3) * 1e3
        res[proc] = entry
    out[key] = res
if rank == 0:
    print(json.dumps(out))
if world > 1:
    from alplib_b200 import peer
    peer.reset()
    dist.destroy_process_group()

#!/usr/bin/env python
"""One rank's share of the C5 Compton scan (masses r::8): back-to-back device time vs latency of a single synchronous call."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
import bench
from alplib_b200 import scan as S
from alplib_b200.runtime import Runtime

W = A.Material("W")
ma, g = np.logspace(-2, 2, 200), np.logspace(-9, -3, 200)
kw = dict(days_exposure=100, threshold=5.0, seed=bench.SEED, **bench.GEOM)
pf = bench.c2_spectrum(1000)
rt = Runtime.get()
for share in (1, 8):
    for r in ((0,) if share == 1 else (0, 3, 7)):
        my = ma[r::share]
        f = lambda: S.scan_rates("compton", pf, W, my, g, n_samples=100, return_device=True, g0=float(g[0]), **kw)
        for _ in range(5):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record(); torch.cuda.synchronize()
        b2b = e0.elapsed_time(e1) / 20
        lat, wall = [], []
        for _ in range(20):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record(); d = f(); e1.record()
            h = rt.to_host(d)
            wall.append((time.perf_counter() - t0) * 1e3)
            lat.append(e0.elapsed_time(e1))
        print(f"share 1/{share} rank {r}: back-to-back {b2b:.3f} ms/call; single call device {np.median(lat):.3f} ms, wall incl. D2H {np.median(wall):.3f} ms")

#!/usr/bin/env python
"""Time the C5 (m_a, g) scan: device time of the Compton and Primakoff maps (run on the GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
from alplib_b200.scan import scan_rates
import bench

W = A.Material("W")
pf = bench.c2_spectrum(bench.N_BINS)[::10]
ma, g = np.logspace(-2, 2, 200), np.logspace(-9, -3, 200)
kw = dict(days_exposure=100, threshold=5.0, seed=1, **bench.GEOM)
for proc, extra in (("compton", dict(n_samples=100)), ("primakoff", {})):
    for _ in range(3):
        m = scan_rates(proc, pf, W, ma, g, **extra, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m = scan_rates(proc, pf, W, ma, g, **extra, **kw)
    e1.record(); torch.cuda.synchronize()
    print(f"{os.environ.get('ALPK_LIB_PATH', 'default'):30s} {proc:10s} {e0.elapsed_time(e1) / 5:8.3f} ms/map  nonzero {np.count_nonzero(m)}  sum {m.sum():.12e}")

#!/usr/bin/env python
"""Chain kernel with the fused event-weighted E_a histogram vs chain + separate histogram kernel (run on the GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
from alplib_b200 import reduce as R
import bench

pf = bench.c2_spectrum(bench.N_BINS)
c = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES,
                           seed=bench.SEED, **bench.GEOM)
c.simulate()


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for nb, kind in ((50, "lin"), (200, "log"), (2000, "log")):
    edges = np.linspace(0, 1000, nb + 1) if kind == "lin" else np.logspace(0, 3, nb + 1)
    ed = c._rt.upload(edges)
    h = c._rt.zeros(nb)
    t_fused = timeit(lambda: c.chain(100, 5.0, materialize=False, hist_edges=ed, hist=h))
    t_fused_mat = timeit(lambda: c.chain(100, 5.0, materialize=True, hist_edges=ed, hist=h))

    def separate():
        rate, evw = c.chain(100, 5.0, materialize=True)
        R.histogram(c.device_arrays()["axion_energy"], evw, ed, out=h)
    t_sep = timeit(separate)
    print(f"{nb:5d} {kind} bins: fused reduce-only {t_fused:.3f} ms, fused materialising {t_fused_mat:.3f} ms, chain + hist kernel {t_sep:.3f} ms")

#!/usr/bin/env python
"""Device time of the reduction kernels on the C2 sample set: weighted histogram, sum, decays (run on the GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
from alplib_b200 import reduce as R
import bench

pf = bench.c2_spectrum(bench.N_BINS)
c = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES,
                           seed=bench.SEED, fused=False, **bench.GEOM)
c.simulate(); c.propagate()
d = c.device_arrays()
E_, F_, D32 = d["axion_energy"], d["axion_flux"], d["decay_axion_weight"]
n = E_.shape[0]


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
    out = c._rt.zeros(nb)
    ms = timeit(lambda: R.histogram(E_, F_, ed, out=out))
    ms32 = timeit(lambda: R.histogram(E_, D32, ed, out=out))
    print(f"hist {nb:5d} {kind} bins: f64 weights {ms:.4f} ms ({n * 16 / ms / 1e6:.0f} GB/s)  f32 weights {ms32:.4f} ms ({n * 12 / ms32 / 1e6:.0f} GB/s)")
ms = timeit(lambda: R.total(F_))
print(f"sum: {ms:.4f} ms ({n * 8 / ms / 1e6:.0f} GB/s)")

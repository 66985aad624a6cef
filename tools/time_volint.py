#!/usr/bin/env python
"""propagate_iso_vol_int throughput: n samples x 1e4 detector-volume points (run on the GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
import bench

pf = bench.c2_spectrum(10000)
fl = A.FluxPrimakoffIsotropic(pf, A.Material("W"), axion_mass=0.1, axion_coupling=1e-5, **bench.GEOM)
fl.simulate()
np.random.seed(4)
geom = A.DetectorGeometry(100.0, "box", [2.0, 2.0, 3.0], n_samples=10000)
for _ in range(2):
    fl.propagate_iso_vol_int(geom)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fl.propagate_iso_vol_int(geom)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
n = fl.n_flux() * 10000
print(f"{os.environ.get('ALPK_PRECISION', 'fast'):9s} vol_int {fl.n_flux()} samples x 1e4 points = {n:.2e} evaluations: {ms:.3f} ms  {n / ms / 1e6:.1f} G eval/s  sum {np.sum(fl.decay_axion_weight, dtype=np.float64):.10e}")

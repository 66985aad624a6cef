#!/usr/bin/env python
"""Device time of the detection-rate kernels (inverse Compton / inverse Primakoff event weights) on the C2 sample set."""
import ctypes as C, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
from alplib_b200 import params as P, _native as N
import bench

pf = bench.c2_spectrum(bench.N_BINS)
c = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES,
                           seed=bench.SEED, fused=False, **bench.GEOM)
c.simulate(); c.propagate()
rt = c._rt
n = c.n_flux()
ev = P.event_params(100, 5.0)
rate = rt.empty(1)
for kind, xs in ((0, P.icompton_consts(1.5, 1e-7, 18)), (2, P.icompton_consts(1.5, 1e-7, 18)),
                 (1, P.iprimakoff_consts(1.5, 1e-7, 18)), (3, P.iprimakoff_consts(1.5, 1e-7, 18))):
    def go():
        N.call("alpk_scatter_events", kind, N.ptr(c._E), N.ptr(c._S32), n, xs, float(1e27 / 0.04), float(P.METER_BY_MEV ** 2),
               C.byref(ev), None, N.ptr(rate), N.ptr(rt.scratch), rt.stream())
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"kind {kind}: {ms:.4f} ms  {n / ms / 1e6:.1f} G samples/s  {n * 12 / ms / 1e6:.0f} GB/s  rate {rate.item():.12e}")

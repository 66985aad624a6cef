#!/usr/bin/env python
"""Launches each hot kernel exactly twice at BASELINE size (warm-up + capture) -- the ncu target:
   ncu --set full -k regex:"chain_rows_kernel|decay2body_kernel|scan_kernel" --launch-skip 5 --launch-count 5 python tools/ncu_target.py
order: chain materialise, chain reduce-only, chain reduce-only + fused 50-bin histogram, decay2body, compton scan."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
from alplib_b200 import params as P
from alplib_b200.scan import scan_rates
import bench

W, Ar = A.Material("W"), A.Material("Ar")
pf = bench.c2_spectrum(bench.N_BINS)
c = A.FluxComptonIsotropic(pf, W, axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES, seed=bench.SEED, **bench.GEOM)
c._stage_inputs(); c._launch_rows()
p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=bench.MA_P, axion_coupling=bench.G_P, **bench.GEOM)
p.simulate(); p.propagate()
g = A.PhotonEventGenerator(p, Ar)
d_edges = c._rt.upload(bench.HIST_EDGES)
d_hist = c._rt.zeros(bench.HIST_EDGES.shape[0] - 1)
for rep in range(2):
    c.chain(100, 5.0, materialize=True)
    c.chain(100, 5.0, materialize=False)
    c.chain(100, 5.0, materialize=False, hist_edges=d_edges, hist=d_hist)
    g.simulate_decay_4vectors(100, 10000, seed=3)
    scan_rates("compton", pf[::10], W, np.logspace(-2, 2, 200), np.logspace(-9, -3, 200), n_samples=100, days_exposure=100,
               threshold=5.0, seed=1, **bench.GEOM)
    torch.cuda.synchronize()
print("ncu target done")

#!/usr/bin/env python
"""cProfile of the public-API (e2e) pass of bench.py: where the host time of one pass goes (run on the GPU box)."""
import cProfile, os, pstats, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
import bench

W, Ar = A.Material("W"), A.Material("Ar")
pf = bench.c2_spectrum(bench.N_BINS)


def e2e_pass():
    c = A.FluxComptonIsotropic(pf, W, axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES, seed=bench.SEED, **bench.GEOM)
    c.simulate(); c.propagate()
    gen_c = A.ElectronEventGenerator(c, Ar)
    rate_c = gen_c.decays_device(bench.DAYS, bench.THR)
    p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=bench.MA_P, axion_coupling=bench.G_P, **bench.GEOM)
    p.simulate(); p.propagate()
    rp = A.PhotonEventGenerator(p, Ar).decays(bench.DAYS, bench.THR)
    return rp, float(rate_c.item())


for _ in range(64):
    e2e_pass()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(256):
    e2e_pass()
dt = (time.perf_counter() - t0) / 256
print(f"e2e pass: {dt * 1e3:.4f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(256):
    e2e_pass()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)

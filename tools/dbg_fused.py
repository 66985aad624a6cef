import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
import bench
W, Ar = A.Material("W"), A.Material("Ar")
pf = bench.c2_spectrum(bench.N_BINS)
def step(fused):
    p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=bench.MA_P, axion_coupling=bench.G_P, fused=fused, **bench.GEOM)
    p.simulate(); p.propagate()
    rp = A.PhotonEventGenerator(p, Ar).decays(100, 5.0)
    c = A.FluxComptonIsotropic(pf, W, axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES, seed=bench.SEED, fused=fused, **bench.GEOM)
    c.simulate(); c.propagate()
    rc = A.ElectronEventGenerator(c, Ar).decays(100, 5.0)
    return rp, rc, p, c
for mode in (False, True, True):
    ts = []
    for i in range(40):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = step(mode)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    print("fused" if mode else "eager", " ".join(f"{x:.2f}" for x in ts))

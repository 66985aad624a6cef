#!/usr/bin/env python
"""Host-side phase timings of the public-API (e2e) step, eager vs fused (run on the GPU box)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
import bench

W, Ar = A.Material("W"), A.Material("Ar")
pf = bench.c2_spectrum(bench.N_BINS)


def run(fused, n=10):
    acc = {}
    def tick(k, t0):
        t1 = time.perf_counter(); acc[k] = acc.get(k, 0.0) + (t1 - t0); return t1
    for it in range(n + 2):
        if it == 2:
            acc.clear()
        torch.cuda.synchronize()
        t = time.perf_counter()
        p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=bench.MA_P, axion_coupling=bench.G_P, fused=fused, **bench.GEOM); t = tick("P.ctor", t)
        p.simulate(); t = tick("P.simulate", t)
        p.propagate(); t = tick("P.propagate", t)
        g = A.PhotonEventGenerator(p, Ar); r = g.decays(100, 5.0); t = tick("P.decays", t)
        c = A.FluxComptonIsotropic(pf, W, axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES, seed=bench.SEED, fused=fused, **bench.GEOM); t = tick("C.ctor", t)
        c.simulate(); t = tick("C.simulate", t)
        c.propagate(); t = tick("C.propagate", t)
        g = A.ElectronEventGenerator(c, Ar); t = tick("C.gen", t)
        r = g.decays(100, 5.0); t = tick("C.decays", t)
        torch.cuda.synchronize(); t = tick("sync", t)
    tot = sum(acc.values()) / n * 1e3
    print("fused" if fused else "eager", f"total {tot:.3f} ms/step :", {k: round(v / n * 1e3, 3) for k, v in acc.items()})


run(False); run(True); run(False); run(True)

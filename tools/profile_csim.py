#!/usr/bin/env python
"""Fine-grained host timings of FluxComptonIsotropic.simulate() pieces (run on the GPU box)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
from alplib_b200 import params as P
import bench

W, Ar = A.Material("W"), A.Material("Ar")
pf = bench.c2_spectrum(bench.N_BINS)
acc = {}
def tick(k, t0):
    t1 = time.perf_counter(); acc[k] = acc.get(k, 0.0) + (t1 - t0); return t1
n = 50
for it in range(n + 5):
    if it == 5:
        acc.clear()
    torch.cuda.synchronize()
    t = time.perf_counter()
    c = A.FluxComptonIsotropic(pf, W, axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES, seed=bench.SEED, **bench.GEOM); t = tick("ctor", t)
    c._reset_samples(); t = tick("reset", t)
    c._stage_inputs(); t = tick("stage", t)
    c._launch_rows(); t = tick("rows", t)
    c._finish_simulate(); t = tick("finish", t)
    c.propagate(); t = tick("propagate", t)
    g = A.ElectronEventGenerator(c, Ar); t = tick("gen", t)
    r = g.decays_device(100, 5.0); t = tick("decays_device", t)
    x = float(r.item()); t = tick("item", t)
print({k: round(v / n * 1e3, 4) for k, v in acc.items()}, "ms")
# pieces of stage
acc.clear()
rt = c._rt
for it in range(n):
    t = time.perf_counter()
    s = 2 * 0.511 * pf[:, 0] + 0.511 ** 2; keep = ~(s < (0.511 + 1.5) ** 2); t = tick("np.filter", t)
    ids = (np.nonzero(keep)[0]).astype(np.uint32); t = tick("np.ids", t)
    a0 = pf[keep, 0]; a1 = pf[keep, 1]; t = tick("np.take", t)
    d0 = rt.upload(a0); t = tick("upload0", t)
    d1 = rt.upload(a1); t = tick("upload1", t)
    d2 = rt.upload(ids, np.uint32); t = tick("upload2", t)
    tb = rt.empty(ids.shape[0] * 12); t = tick("empty", t)
    par = P.compton_params(1.5, 1e-7, 74, 10000, fast=True); t = tick("params", t)
    le, lx, nt = c.target_photon_xs.device_table(rt); t = tick("devtable", t)
    st = rt.stream(); t = tick("stream", t)
    torch.cuda.synchronize()
print({k: round(v / n * 1e3, 4) for k, v in acc.items()}, "ms")

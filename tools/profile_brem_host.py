import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import alplib_b200 as A
Cm, Ar = A.Material("C"), A.Material("Ar")
E = np.linspace(20, 1.2e5, 10_000)
ele = np.stack([E, 1e14 * np.exp(-E / 2e4)], axis=1)
pos = np.stack([E, 0.3 * 1e14 * np.exp(-E / 2e4)], axis=1)
DUNE = dict(det_dist=574.0, det_length=5.0, det_area=21.0)
f = A.FluxBremIsotropic(ele, pos, Cm, target_length=100, axion_mass=10.0, axion_coupling=1e-6, n_samples=1000, seed=5, **DUNE)
acc = {}
for it in range(30):
    if it == 10: acc = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    f = A.FluxBremIsotropic(ele, pos, Cm, target_length=100, axion_mass=10.0, axion_coupling=1e-6, n_samples=1000, seed=5, **DUNE)
    t1 = time.perf_counter(); f.simulate(); t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
    f.propagate(); g = A.ElectronEventGenerator(f, Ar); r = g.decays(100, 5.0); t4 = time.perf_counter()
    for k, v in (("ctor", t1 - t0), ("simulate", t2 - t1), ("sync after simulate", t3 - t2), ("propagate+decays", t4 - t3)):
        acc[k] = acc.get(k, 0) + v
print({k: round(v / 20 * 1e3, 4) for k, v in acc.items()})
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    f = A.FluxBremIsotropic(ele, pos, Cm, target_length=100, axion_mass=10.0, axion_coupling=1e-6, n_samples=1000, seed=5, **DUNE)
    f.simulate()
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(14)

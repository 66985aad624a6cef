#!/usr/bin/env python
"""cProfile of the public-API steps of BASELINE config 4 (brem / resonance / pair annihilation): host-side hot spots."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
Cm, Ar = A.Material("C"), A.Material("Ar")
E = np.linspace(20, 1.2e5, 10_000)
ele = np.stack([E, 1e14 * np.exp(-E / 2e4)], axis=1)
pos = np.stack([E, 0.3 * 1e14 * np.exp(-E / 2e4)], axis=1)
DUNE = dict(det_dist=574.0, det_length=5.0, det_area=21.0)
KW = dict(target_length=100, axion_mass=10.0, axion_coupling=1e-6, **DUNE)


def brem():
    f = A.FluxBremIsotropic(ele, pos, Cm, n_samples=1000, seed=5, **KW)
    f.simulate(); f.propagate()
    return A.ElectronEventGenerator(f, Ar).decays(100, 5.0)


def res():
    f = A.FluxResonanceIsotropic(pos, Cm, n_samples=1_000_000, seed=6, **KW)
    f.simulate(); f.propagate()
    return A.ElectronEventGenerator(f, Ar).decays(100, 5.0)


def pair():
    f = A.FluxPairAnnihilationIsotropic(pos, Cm, n_samples=100, seed=7, **KW)
    f.simulate(); f.propagate()
    return A.ElectronEventGenerator(f, Ar).decays(100, 5.0)


for name, fn in (("brem", brem), ("resonance", res), ("pairann", pair)):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        fn()
    print(f"== {name}: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per step")
    pr = cProfile.Profile(); pr.enable()
    for _ in range(50):
        fn()
    pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(9)

#!/usr/bin/env python
"""Throughput of the BASELINE config-4 channels (brems / resonance / pair annihilation on graphite, DUNE-like geometry)
through the public classes: simulate + propagate + decays, device time (run on the GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A

Cm, Ar = A.Material("C"), A.Material("Ar")
E = np.linspace(20, 1.2e5, 10_000)
ele = np.stack([E, 1e14 * np.exp(-E / 2e4)], axis=1)
pos = np.stack([E, 0.3 * 1e14 * np.exp(-E / 2e4)], axis=1)
DUNE = dict(det_dist=574.0, det_length=5.0, det_area=21.0)


def timeit(fn, n=20, warm=8):
    for _ in range(warm):
        r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


def brem():
    f = A.FluxBremIsotropic(ele, pos, Cm, target_length=100, axion_mass=10.0, axion_coupling=1e-6, n_samples=1000, seed=5, **DUNE)
    f.simulate(); f.propagate()
    return A.ElectronEventGenerator(f, Ar).decays(100, 5.0), f.n_flux()


def res():
    f = A.FluxResonanceIsotropic(pos, Cm, target_length=100, axion_mass=10.0, axion_coupling=1e-6, n_samples=1_000_000, seed=6, **DUNE)
    f.simulate(); f.propagate()
    return A.ElectronEventGenerator(f, Ar).decays(100, 5.0), 1_000_000


def pair():
    f = A.FluxPairAnnihilationIsotropic(pos, Cm, target_length=100, axion_mass=10.0, axion_coupling=1e-6, n_samples=100, seed=7, **DUNE)
    f.simulate(); f.propagate()
    return A.ElectronEventGenerator(f, Ar).decays(100, 5.0), f.n_flux()


for name, fn in (("brem", brem), ("resonance", res), ("pairann", pair)):
    ms, (rate, n) = timeit(fn)
    print(f"{name:10s} {n:.3e} samples  {ms:8.3f} ms/step (API, end to end)  {n / ms / 1e6:8.3f} G samples/s  rate {rate:.6e}")

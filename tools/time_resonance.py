#!/usr/bin/env python
"""FluxResonanceIsotropic.simulate() device time vs inner sample count (run on the GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A

Cm = A.Material("C")
E = np.linspace(20, 1.2e5, 10_000)
pos = np.stack([E, 0.3 * 1e14 * np.exp(-E / 2e4)], axis=1)
DUNE = dict(det_dist=574.0, det_length=5.0, det_area=21.0)
for n in (1_000_000, 10_000_000, 100_000_000):
    f = A.FluxResonanceIsotropic(pos, Cm, target_length=100, axion_mass=10.0, axion_coupling=1e-6, n_samples=n, seed=6, **DUNE)
    for _ in range(2):
        f.simulate()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f.simulate()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{os.environ.get('ALPK_PRECISION', 'fast'):8s} resonance n={n:.0e}: {ms:.3f} ms  {n / ms / 1e6:.2f} G samples/s  flux {np.asarray(f.axion_flux)[0]:.12e}")

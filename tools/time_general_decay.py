#!/usr/bin/env python
"""General two-body decay kernel (dense boost): neutral mesons decaying in flight, 1e4 mesons x 1e4 decays."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A

rng = np.random.default_rng(1)
n_m, ns = 10000, 10000
pz = 10 ** rng.uniform(2, 4, n_m); pt = pz * 10 ** rng.uniform(-3, -1, n_m); ph = rng.uniform(0, 2 * np.pi, n_m)
mf = np.stack([pt * np.cos(ph), pt * np.sin(ph), pz, np.sqrt(pz ** 2 + pt ** 2 + 134.98 ** 2)], axis=1)
fl = A.FluxNeutralMeson2BodyDecay(mf, 1.0, boson_mass=30.0, coupling=1e-3, n_samples=ns, seed=2)
for _ in range(2):
    fl._reset_samples(); fl._A = None
    fl.simulate()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    fl._reset_samples(); fl._A = None
    fl.simulate()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"neutral meson 2-body: {n_m * ns:.1e} decays in {ms:.3f} ms = {n_m * ns / ms / 1e6:.1f} G decays/s; kept {len(fl.axion_energy)}")

#!/usr/bin/env python
"""Device time of the bremsstrahlung / pair-annihilation chain kernels at BASELINE config-4 size (run on the GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
from alplib_b200 import params as P

Cm = A.Material("C")
E = np.linspace(20, 1.2e5, 10_000)
ele = np.stack([E, 1e14 * np.exp(-E / 2e4)], axis=1)
pos = np.stack([E, 0.3 * 1e14 * np.exp(-E / 2e4)], axis=1)
DUNE = dict(det_dist=574.0, det_length=5.0, det_area=21.0)


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, f in (("brem", A.FluxBremIsotropic(ele, pos, Cm, target_length=100, axion_mass=10.0, axion_coupling=1e-6, n_samples=1000, seed=5, **DUNE)),
                ("brem-vector", A.FluxBremIsotropic(ele, pos, Cm, target_length=100, axion_mass=10.0, axion_coupling=1e-6, n_samples=1000, seed=5, boson_type="vector", **DUNE)),
                ("pairann", A.FluxPairAnnihilationIsotropic(pos, Cm, target_length=100, axion_mass=10.0, axion_coupling=1e-6, n_samples=1000, seed=7, **DUNE))):
    f.fused = True
    f.simulate()
    rt = f._rt
    n = f._rows[1] * f._rows[2]
    Ea, Fl = rt.empty(n), rt.empty(n)
    d32, s32, evw, rate = rt.empty(n, torch.float32), rt.empty(n, torch.float32), rt.empty(n), rt.empty(1)
    prop = f._prop_params(*f._width_rescale(None), f._geom_accept(), None)
    ev = P.event_params(100, 5.0)
    t0 = timeit(lambda: f._launch_samples(Ea, Fl, None, None))
    t1 = timeit(lambda: f._launch_samples(Ea, Fl, prop, ev, d32=d32, s32=s32, evw=evw, rate=rate))
    print(f"{os.environ.get('ALPK_PRECISION', 'fast'):8s} {name:12s} n={n:.3e}  prod {t0:.4f} ms ({n / t0 / 1e6:.1f} G/s)  chain_mat {t1:.4f} ms ({n / t1 / 1e6:.1f} G/s)  rate {rate.item():.12e}")

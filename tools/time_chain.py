#!/usr/bin/env python
"""Time the Compton chain kernel variants (ALPK_LIB_PATH selects the library build)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
from alplib_b200 import params as P
import bench


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


W = A.Material("W")
pf = bench.c2_spectrum(bench.N_BINS)
c = A.FluxComptonIsotropic(pf, W, axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES, seed=bench.SEED, **bench.GEOM)
c._stage_inputs(); c._launch_rows()
n = c._rows[1] * c._rows[2]
rt = c._rt
E, F = rt.empty(n), rt.empty(n)
d32, s32, evw, rate = rt.empty(n, torch.float32), rt.empty(n, torch.float32), rt.empty(n), rt.empty(1)
prop = c._prop_params(*c._width_rescale(None), c._geom_accept(), None)
ev = P.event_params(100, 5.0)
t0 = timeit(lambda: c._launch_samples(E, F, None, None))
t1 = timeit(lambda: c._launch_samples(E, F, prop, ev, d32=d32, s32=s32, evw=evw, rate=rate))
t2 = timeit(lambda: c._launch_samples(None, None, prop, ev, rate=rate))
print(f"{os.environ.get('ALPK_LIB_PATH', 'default'):40s} prod {t0:.4f} ms  chain_mat {t1:.4f} ms ({n / t1 / 1e6:.1f} G/s)  reduce_only {t2:.4f} ms  rate {rate.item():.12e}")

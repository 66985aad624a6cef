#!/usr/bin/env python
"""Per-kernel CUDA-event timings of the hot path at BASELINE sizes (run on the GPU box; not part of the tests)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A                       # noqa: E402
from alplib_b200 import params as P           # noqa: E402
import bench                                   # noqa: E402


def timeit(fn, n=10, warm=3):
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


def main():
    out = {}
    W, Ar = A.Material("W"), A.Material("Ar")
    pf = bench.c2_spectrum(bench.N_BINS)
    c = A.FluxComptonIsotropic(pf, W, axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES,
                               seed=bench.SEED, **bench.GEOM)
    c._stage_inputs(); c._launch_rows()
    n = c._rows[1] * c._rows[2]
    out["n_compton"] = n
    rt = c._rt
    E, F = rt.empty(n), rt.empty(n)
    d32, s32, evw, rate = rt.empty(n, torch.float32), rt.empty(n, torch.float32), rt.empty(n), rt.empty(1)
    prop = c._prop_params(*c._width_rescale(None), c._geom_accept(), None)
    ev = P.event_params(100, 5.0)
    out["rows_ms"] = timeit(c._launch_rows)
    out["prod_only_ms"] = timeit(lambda: c._launch_samples(E, F, None, None))
    out["prod_prop_ms"] = timeit(lambda: c._launch_samples(E, F, prop, None, d32=d32, s32=s32))
    out["chain_materialize_ms"] = timeit(lambda: c._launch_samples(E, F, prop, ev, d32=d32, s32=s32, evw=evw, rate=rate))
    out["chain_reduce_only_ms"] = timeit(lambda: c._launch_samples(None, None, prop, ev, rate=rate))
    c._set_production(E, F)
    out["propagate_ms"] = timeit(lambda: c._run_propagate(prop))
    gen = A.ElectronEventGenerator(c, Ar)
    out["decays_ms"] = timeit(lambda: gen.decays_device(100, 5.0))
    gen.materialize_weights = False
    out["compton_scatter_ms"] = timeit(lambda: gen._scatter(0, P.icompton_consts(1.5, 1e-7, 18), 1e27, 100, 5.0))
    # decay MC: 1e4 parents x 1e4 decays
    p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=bench.MA_P, axion_coupling=bench.G_P, **bench.GEOM)
    p.simulate(); p.propagate()
    g = A.PhotonEventGenerator(p, Ar)
    out["n_decay_events"] = p.n_flux() * 10000
    out["decay4_ms"] = timeit(lambda: g.simulate_decay_4vectors(100, 10000, seed=3), n=5, warm=2)
    # host-side cost of one e2e step
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        cc = A.FluxComptonIsotropic(pf, W, axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES,
                                    seed=bench.SEED, fused=True, **bench.GEOM)
        t1 = time.perf_counter(); cc.simulate(); t2 = time.perf_counter(); cc.propagate(); t3 = time.perf_counter()
        r = A.ElectronEventGenerator(cc, Ar).decays(100, 5.0); t4 = time.perf_counter()
    out["e2e_fused_host_ms"] = {"ctor": (t1 - t) * 1e3 / 5, "simulate": (t2 - t1) * 1e3, "propagate": (t3 - t2) * 1e3,
                                "decays": (t4 - t3) * 1e3}
    for k, v in out.items():
        if k.endswith("_ms") and not isinstance(v, dict):
            nn = out["n_decay_events"] if k == "decay4_ms" else n
            print(f"{k:28s} {v:9.4f} ms   {nn / (v * 1e-3):.3e} /s")
        else:
            print(k, v)
    json.dump(out, open(os.path.join("gpurun_out", "profile_steps.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

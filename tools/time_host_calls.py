#!/usr/bin/env python
"""Host-side cost (microseconds per call, GPU work asynchronous) of the pieces of one public-API Compton / Primakoff job."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
import bench
from alplib_b200.constants import M_E

W, Ar = A.Material("W"), A.Material("Ar")
pf = bench.c2_spectrum(bench.N_BINS)
rt = A.runtime.Runtime.get() if hasattr(A, "runtime") else None
from alplib_b200.runtime import Runtime
rt = Runtime.get()


def t(label, fn, n=300, sync_every=20):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(n):
        t0 = time.perf_counter()
        fn()
        tot += time.perf_counter() - t0
        if i % sync_every == sync_every - 1:
            torch.cuda.synchronize()
    print(f"{label:46s} {tot / n * 1e6:8.1f} us")


t("rt.stream()", rt.stream)
t("rt.empty(1)", lambda: rt.empty(1))
t("rt.empty_block(120000)", lambda: rt.empty_block(120000))
t("rt.stage_flux_rows (1e4 rows)", lambda: rt.stage_flux_rows(pf, 1, float((M_E + 1.5) ** 2), 0))
t("Compton ctor", lambda: A.FluxComptonIsotropic(pf, W, axion_mass=1.5, axion_coupling=1e-7, n_samples=10000, seed=1, **bench.GEOM))
c = A.FluxComptonIsotropic(pf, W, axion_mass=1.5, axion_coupling=1e-7, n_samples=10000, seed=1, **bench.GEOM)
t("Compton _stage_inputs", c._stage_inputs)
t("Compton _launch_rows", c._launch_rows)
t("Compton simulate()", c.simulate)
t("Compton propagate()", c.propagate)
g = A.ElectronEventGenerator(c, Ar)
t("ElectronEventGenerator ctor", lambda: A.ElectronEventGenerator(c, Ar))


def dd():
    c.simulate(); c.propagate()
    return g.decays_device(100, 5.0)


t("Compton simulate+propagate+decays_device", dd, n=100, sync_every=1)
p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=0.1, axion_coupling=1e-5, **bench.GEOM)
t("Primakoff ctor", lambda: A.FluxPrimakoffIsotropic(pf, W, axion_mass=0.1, axion_coupling=1e-5, **bench.GEOM))
t("Primakoff simulate()", p.simulate)
t("Primakoff propagate()", p.propagate)
gp = A.PhotonEventGenerator(p, Ar)


def pd():
    p.simulate(); p.propagate()
    return gp.decays_device(100, 5.0)


t("Primakoff simulate+propagate+decays_device", pd, n=300, sync_every=5)
t("PendingValue(rate)", lambda: A.runtime.PendingValue(rt.empty(1)) if hasattr(A, "runtime") else None)

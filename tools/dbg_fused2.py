import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
import bench
W, Ar = A.Material("W"), A.Material("Ar")
pf = bench.c2_spectrum(bench.N_BINS)
names = ["P.ctor", "P.sim", "P.prop", "P.dec", "C.ctor", "C.sim", "C.prop", "C.dec"]
def step(fused, sync):
    ts = []
    def tk():
        if sync: torch.cuda.synchronize()
        ts.append(time.perf_counter())
    tk()
    p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=bench.MA_P, axion_coupling=bench.G_P, fused=fused, **bench.GEOM); tk()
    p.simulate(); tk(); p.propagate(); tk()
    rp = A.PhotonEventGenerator(p, Ar).decays(100, 5.0); tk()
    c = A.FluxComptonIsotropic(pf, W, axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES, seed=bench.SEED, fused=fused, **bench.GEOM); tk()
    c.simulate(); tk(); c.propagate(); tk()
    rc = A.ElectronEventGenerator(c, Ar).decays(100, 5.0); tk()
    return [(b - a) * 1e3 for a, b in zip(ts[:-1], ts[1:])]
for _ in range(10): step(False, False)
import gc
for sync in (False, True):
    for gcoff in (False, True):
        if gcoff: gc.disable()
        else: gc.enable()
        rows = [step(True, sync) for _ in range(40)]
        tot = [sum(r) for r in rows]
        slow = [i for i, t in enumerate(tot) if t > 1.5]
        print(f"sync={sync} gc_off={gcoff} totals:", " ".join(f"{t:.2f}" for t in tot))
        for i in slow[:3]:
            print("   slow step", i, {n: round(v, 2) for n, v in zip(names, rows[i])})

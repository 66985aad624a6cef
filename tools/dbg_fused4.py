import os, sys, time, gc
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A
import bench
W, Ar = A.Material("W"), A.Material("Ar")
pf = bench.c2_spectrum(bench.N_BINS)
def step(fused):
    c = A.FluxComptonIsotropic(pf, W, axion_mass=bench.MA_C, axion_coupling=bench.G_C, n_samples=bench.N_SAMPLES, seed=bench.SEED, fused=fused, **bench.GEOM)
    c.simulate(); c.propagate()
    rate_c = A.ElectronEventGenerator(c, Ar).decays_device(100, 5.0)
    p = A.FluxPrimakoffIsotropic(pf, W, axion_mass=bench.MA_P, axion_coupling=bench.G_P, fused=fused, **bench.GEOM)
    p.simulate(); p.propagate()
    rp = A.PhotonEventGenerator(p, Ar).decays(100, 5.0)
    rc = float(rate_c.item())
    return rp, rc, p, c
def run(tag, fused, n, pre=None):
    if pre: pre()
    ts = []
    for i in range(n):
        t = time.perf_counter(); r = step(fused); ts.append((time.perf_counter() - t) * 1e3)
    big = [(i, round(x, 2)) for i, x in enumerate(ts) if x > 1.3]
    print(f"{tag:34s} median {np.median(ts):.3f} max {max(ts):.2f}  slow steps: {big[:12]}")
mode = sys.argv[1] if len(sys.argv) > 1 else "a"
run("auto warm", None, 20)
run("eager", False, 20)
if mode == "a":
    run("fused after empty_cache", True, 120, torch.cuda.empty_cache)
elif mode == "b":
    run("fused no empty_cache", True, 120)
elif mode == "c":
    gc.disable(); run("fused no empty_cache, gc off", True, 120)
elif mode == "d":
    gc.freeze(); run("fused after empty_cache, gc frozen", True, 120, torch.cuda.empty_cache)

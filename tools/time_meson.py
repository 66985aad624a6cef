#!/usr/bin/env python
"""Throughput of the charged-meson 3-body kernels (run on the GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alplib_b200 as A

n_m, ns = 5000, 1000
for model in ("vector_ib1", "scalar_ib1", "pseudoscalar_ib1", "vector_ib9", "vector_contact", "scalar_ib1"):
    mf = np.stack([np.linspace(10.0, 5000.0, n_m), np.full(n_m, 1.0)], axis=1)
    fl = A.FluxChargedMeson3BodyIsotropic(meson_flux=mf, boson_mass=5.0, coupling=1e-3, meson_type="kaon",
                                          interaction_model=model, n_samples=ns, seed=3, abd=(1.0, 0.5, 0.2))
    for _ in range(2):
        fl.simulate()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fl.simulate()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    n = n_m * 2 * ns
    print(f"iso   {model:16s} {n:.2e} samples  {ms:8.3f} ms  {n / ms / 1e6:8.2f} G samples/s  sum {np.sum(fl.axion_flux):.10e}")
mfd = np.stack([np.linspace(100.0, 5000.0, n_m), np.full(n_m, 0.01), np.full(n_m, 1.0)], axis=1)
fd = A.FluxChargedMeson3BodyDecay(mfd, boson_mass=5.0, coupling=1e-3, n_samples=ns, meson_type="kaon", seed=4)
for _ in range(2):
    fd.simulate()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    fd.simulate()
e1.record(); torch.cuda.synchronize()
print(f"decay scalar_ib1 {n_m * 2 * ns:.2e} samples {e0.elapsed_time(e1) / 3:8.3f} ms  kept {len(fd.axion_energy)}  sum {np.sum(fd.axion_flux):.10e}")

#!/usr/bin/env python
"""(m_a, g) sensitivity scan of a beam-dump photon spectrum: Primakoff + Compton production, propagation and decay event
rates on a 200 x 200 grid in two kernel launches, then two-sided limits per mass (reference pattern: README.md:199-226,
fit.py:12-58 -- one full simulate/propagate/decays chain per grid point).  Needs a B200; `torchrun --nproc-per-node N`
shards the mass axis over N GPUs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))   # repo root
import torch.distributed as dist

from alplib_b200 import Material
from alplib_b200.fit import cleanLimitData, limits_from_rate_map
from alplib_b200.scan import scan_rates, scan_rates_distributed

e = np.logspace(np.log10(0.5), 3, 1000)
photons = np.stack([e, 1e16 * e ** -1.5 * np.exp(-e / 150)], axis=1)          # photons / s per bin in the W dump
ma_grid, g_grid = np.logspace(-2, 2, 200), np.logspace(-9, -3, 200)
kw = dict(det_dist=4.0, det_length=0.2, det_area=0.04, days_exposure=100, threshold=5.0, seed=1)

def main():
    """Returns (compton map, primakoff map, {name: (masses, limits)}); under torchrun every rank gets the full maps."""
    distributed = "RANK" in os.environ
    if distributed:
        import torch
        world, local = int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
        n_dev = torch.cuda.device_count()
        torch.cuda.set_device(local % n_dev)
        # one GPU per rank -> NCCL over NVLink; fewer GPUs than ranks (a 1-GPU test box) -> the ranks share devices and the
        # small merge goes through gloo
        dist.init_process_group("nccl" if n_dev >= world else "gloo")
        run = scan_rates_distributed
    else:
        run = scan_rates
    compton = run("compton", photons, Material("W"), ma_grid, g_grid, n_samples=100, **kw)
    primakoff = run("primakoff", photons, Material("W"), ma_grid, g_grid, **kw)
    limits = {}
    for name, rates in (("electron coupling (Compton)", compton), ("photon coupling (Primakoff)", primakoff)):
        lo, hi = limits_from_rate_map(g_grid, rates, stop_value=2.71)
        ok = ~np.isnan(lo)
        limits[name] = cleanLimitData(ma_grid[ok], lo[ok], hi[ok])
        if not distributed or dist.get_rank() == 0:
            print(f"{name}: {ok.sum()} masses excluded; at ma = {ma_grid[ok][0]:.3g} MeV: {lo[ok][0]:.3g} < g < {hi[ok][0]:.3g}")
    if distributed:
        dist.destroy_process_group()
    return compton, primakoff, limits


if __name__ == "__main__":
    main()

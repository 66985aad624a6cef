#!/usr/bin/env python
"""One beam-dump job on several GPUs: Primakoff + Compton ALP production from a photon spectrum in a tungsten dump, decay event
rates and the event-weighted E_a spectrum in an argon detector (reference pattern: the per-bin loops of fluxes.py:150-151,
226-233 followed by generators.py:48-52 and a np.histogram of the weights).

    python examples/ex_sharded_beam_dump.py                               # one GPU
    torchrun --nproc-per-node 8 examples/ex_sharded_beam_dump.py          # the photon rows sharded over 8 GPUs

Every rank takes an equal block of the photon rows that survive the production threshold (dist.shard_flux), runs the one-kernel
chains, and the ranks' [rates || spectrum] are merged by ONE kernel per rank over NVLink peer memory
(dist.merged_rates_and_hists).  The result does not depend on the number of ranks: the Philox variates are keyed by the global
row index."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))   # repo root
import torch
import torch.distributed as dist

import alplib_b200 as A
from alplib_b200 import dist as D

e = np.logspace(np.log10(0.5), 3, 2000)
photons = np.stack([e, 1e16 * e ** -1.5 * np.exp(-e / 150)], axis=1)          # photons / s per bin in the W dump
geom = dict(det_dist=4.0, det_length=0.2, det_area=0.04)
edges = np.logspace(np.log10(1.5), 3, 41)                                       # E_a bins of the decay spectrum [MeV]


def main(n_samples=2000, seed=7):
    """Returns (Compton decay rate, Primakoff decay rate, event-weighted E_a spectrum); every rank gets the merged result."""
    distributed = "RANK" in os.environ
    if distributed:
        world, local = int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
        n_dev = torch.cuda.device_count()
        torch.cuda.set_device(local % n_dev)
        dist.init_process_group("nccl" if n_dev >= world else "gloo")
    W, Ar = A.Material("W"), A.Material("Ar")
    compton = A.FluxComptonIsotropic(photons, W, axion_mass=1.5, axion_coupling=1e-7, n_samples=n_samples, seed=seed, **geom)
    primakoff = A.FluxPrimakoffIsotropic(photons, W, axion_mass=0.1, axion_coupling=1e-5, **geom)
    if distributed:
        D.shard_flux(compton)                   # this rank's block of kept rows; row_offset keeps the global Philox keys
        D.shard_flux(primakoff)
    compton.simulate()
    primakoff.simulate()
    rate_c, _ = compton.chain(100, 5.0, materialize=False, hist_edges=edges)   # production -> propagate -> decays + spectrum
    rate_p, _ = primakoff.chain(100, 5.0, materialize=False)
    rates, (spectrum,) = D.merged_rates_and_hists([rate_c, rate_p], [compton.last_hist])
    rates, spectrum = rates.cpu().numpy(), spectrum.cpu().numpy()
    if not distributed or dist.get_rank() == 0:
        print(f"decay events in 100 days above 5 MeV: Compton-produced {rates[0]:.6g}, Primakoff-produced {rates[1]:.6g}; "
              f"spectrum peaks at E_a = {edges[np.argmax(spectrum)]:.3g} MeV")
    if distributed:
        from alplib_b200 import peer
        peer.reset()
        dist.destroy_process_group()
    return rates[0], rates[1], spectrum


if __name__ == "__main__":
    main()

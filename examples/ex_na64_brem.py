#!/usr/bin/env python
"""NA64 missing-energy style flux from ALP bremsstrahlung (reference: examples/ex4_na64.py, README.md:176-236) with the
B200 drop-in.  The reference script omits `positron_flux`, which makes its default track-length mode raise TypeError
(SURVEY.md section 4 item 2); a monoenergetic beam is expressed with is_monoenergetic=True.  Needs a B200."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))   # repo root

from alplib_b200 import *   # noqa: F401,F403

beam_energy = 1e5  # 100 GeV
eot = 2.84e11
days = 365
na64_dump = Material("W")
na64_ecal = Material("Pb")
ecal_length, ecal_dist, ecal_area = 1.0, 43.3, 0.23
dump_length = 2.0
ecal_thresh = 1.0e3
na64_e_flux = np.array([[beam_energy, eot / S_PER_DAY / days]])
ecal_n = 100.0 * MEV_PER_KG / na64_ecal.m[0]


def get_events(ma, g, use_loop=False):
    flux_brem = FluxBremIsotropic(na64_e_flux, target=na64_dump, det_dist=ecal_dist, det_length=ecal_length,
                                  det_area=ecal_area, target_length=dump_length, axion_mass=ma, axion_coupling=g,
                                  loop_decay=use_loop, is_isotropic=False, n_samples=10000, is_monoenergetic=True)
    flux_brem.simulate()
    flux_brem.propagate()
    events_gen = ElectronEventGenerator(flux_brem, na64_ecal)
    n_compton = events_gen.compton(ge=g, ma=ma, ntargets=ecal_n, days_exposure=days, threshold=ecal_thresh)
    n_decay = events_gen.decays(days_exposure=days, threshold=ecal_thresh)
    return events_gen.axion_energy, events_gen.decay_weights + events_gen.scatter_weights, n_decay, n_compton


if __name__ == "__main__":
    for ma, g in ((0.5, 1e-5), (10.0, 1e-6), (50.0, 3e-5)):
        energies, weights, n_decay, n_compton = get_events(ma, g, use_loop=True)
        hist, edges = np.histogram(1e-3 * energies, weights=weights, bins=20)
        print(f"ma = {ma} MeV, g = {g}: decays = {n_decay:.4g}, inverse-Compton = {n_compton:.4g}, "
              f"spectrum peak at {edges[np.argmax(hist)]:.1f} GeV")

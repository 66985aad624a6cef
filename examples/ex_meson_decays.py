#!/usr/bin/env python
"""Dark bosons from charged-meson 3-body decays M -> l nu X (the reference's FluxChargedMeson3BodyIsotropic /
FluxChargedMeson3BodyDecay, fluxes.py:605-922): a stopped-pion source (isotropic) and a focused kaon beam decaying in
flight towards a far detector.  Needs a B200."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))   # repo root

from alplib_b200 import *   # noqa: F401,F403  (was: from alplib.fluxes import *)

# stopped / slow pions: rows of [momentum (MeV), pions per second]
pions = np.array([[0.0, 1e15], [30.0, 4e14], [80.0, 1e14]])
iso = FluxChargedMeson3BodyIsotropic(meson_flux=pions, boson_mass=20.0, coupling=1e-4, meson_type="pion",
                                     interaction_model="scalar_ib1", det_dist=20.0, det_length=2.0, det_area=2.0,
                                     n_samples=20000, seed=1)
iso.simulate()
iso.propagate()
print(f"isotropic pion source: {len(iso.axion_energy)} samples, total boson rate {np.sum(iso.axion_flux):.4e} /s, "
      f"through the detector {np.sum(iso.scatter_axion_weight):.4e} /s, BR = {iso.total_br():.4e}")

# kaons in flight: rows of [momentum (MeV), polar angle (rad), kaons per spill]
rng = np.random.default_rng(3)
kaons = np.stack([rng.uniform(2e3, 2e4, 2000), np.abs(rng.normal(0.0, 0.01, 2000)), np.full(2000, 5e9)], axis=1)
beam = FluxChargedMeson3BodyDecay(kaons, boson_mass=50.0, coupling=3e-7, n_samples=500, meson_type="kaon",
                                  interaction_model="vector_ib1", energy_cut=500.0, seed=2)
beam.simulate()
beam.propagate(decay="muon")
E_, w = np.asarray(beam.axion_energy), np.asarray(beam.axion_flux)
print(f"kaon beam: {len(E_)} accepted samples of {2000 * 2 * 500}, mean boson energy {np.average(E_, weights=w):.1f} MeV, "
      f"mean angle to the beam axis {np.average(beam.axion_angle, weights=w):.4f} rad, "
      f"decaying inside the detector {np.sum(beam.decay_axion_weight, dtype=np.float64):.4e}")

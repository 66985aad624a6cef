#!/usr/bin/env python
"""README Primakoff example of the reference (README.md:88-117), unchanged apart from the import and the 2-D photon
table (the README's 1-D `gammas` raises IndexError in the reference as well).  Needs a B200."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))   # repo root

from alplib_b200 import *   # noqa: F401,F403  (was: from alplib.fluxes import *; from alplib.generators import *)

wtarget = Material("W")
gammas = np.array([[100.0, 1.0e12]])  # 100 MeV, 1e12 photons / s

flux_p = FluxPrimakoffIsotropic(photon_flux=gammas, target=wtarget, det_dist=4.0, det_length=0.2,
                                det_area=0.04, axion_mass=0.1, axion_coupling=1e-5, n_samples=1000)
flux_p.simulate()   # production flux
flux_p.propagate()  # propagate to the detector, taking into account decays

gen = PhotonEventGenerator(flux_p, Material("Ar"))
print("decays in 100 days above 5 MeV:", gen.decays(days_exposure=100, threshold=5.0))    # 284.1535949707031
print("per-ALP weights:", gen.decay_weights)

p41, p42, wgt = gen.simulate_decay_4vectors(days_exposure=100, n_samples=10000)
energies1 = np.array([p4.energy() for p4 in p41])      # LorentzVector objects, like the reference ...
angles1 = p41.theta()                                   # ... or whole-array accessors
print("mean photon-1 energy:", energies1.mean(), " mean polar angle:", angles1.mean(), " sum of weights:", np.sum(wgt))

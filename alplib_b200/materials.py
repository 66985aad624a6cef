"""Material container -- host-side input, same attributes and JSON schema as the reference (materials.py:12-46)."""
import json
import os

import numpy as np

from .constants import MEV_PER_KG

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
_MAT_FILE = None


def data_path(*parts):
    return os.path.join(_DATA, *parts)


def _materials():
    global _MAT_FILE
    if _MAT_FILE is None:
        with open(data_path("mat_params.json"), "r") as f:
            _MAT_FILE = json.load(f)
    return _MAT_FILE


class Material:
    """
    Target / detector material (all units MeV, kg, cm unless stated).  Reads data/mat_params.json, which is
    consumed unchanged from the reference.  `ntargets` / `ndensity` use the constructor's `density` argument, not the
    JSON density -- reference behaviour (materials.py:42-43), kept.
    """

    def __init__(self, material_name, fiducial_mass=1.0, volume=1.0, density=1.0, efficiency=None):
        self.mat_name = material_name
        self.efficiency = efficiency
        mat_file = _materials()
        if material_name in mat_file:
            mat_info = mat_file[material_name]
            self.iso = mat_info['iso']
            self.z = np.array(mat_info['z'])
            self.n = np.array(mat_info['n'])
            self.m = np.array(mat_info['m'])
            self.frac = np.array(mat_info['frac'])
            self.lattice_const = np.array(mat_info['lattice_const'])
            self.cell_volume = np.array(mat_info['cell_volume'])
            self.r0 = np.array(mat_info['atomic_radius'])
            self.density = mat_info['density']
            self.fid_mass = fiducial_mass
            self.volume = volume
            molar = np.dot(self.m, self.iso * self.frac) / MEV_PER_KG / 1e-3
            self.ntargets = density * volume / molar
            self.ndensity = density / molar
            self.rad_length = mat_info['rad_length']
        else:
            raise Exception("No such detector in mat_params.json.")

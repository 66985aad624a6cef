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


class DetectorGeometry:
    """Monte-Carlo description of the detector volume used by `propagate_iso_vol_int` (reference:
    materials.py:51-124).  Host-side input container: `n_samples` points drawn with the global NumPy RNG at
    construction, exactly as the reference draws them (three `np.random.uniform(0, 1, n)` calls in this order), and
    mapped to box / sphere / cylinder coordinates -- including the reference's cylinder mapping, which takes the
    azimuth from the third variate (materials.py:82).  The per-sample volume integral itself runs on the GPU."""

    def __init__(self, distance_to_source, geometry="box", params=[1.0, 1.0, 1.0], n_samples=10000):
        if distance_to_source <= 0.0:
            raise Exception("Can't have the detector closer than 0.0 (distance_to_source < 0.0)")
        self.l0 = distance_to_source
        self.n_samples = n_samples
        self.u1 = np.random.uniform(0, 1, n_samples)
        self.u2 = np.random.uniform(0, 1, n_samples)
        self.u3 = np.random.uniform(0, 1, n_samples)
        self.geom = geometry
        self.params = params
        if geometry == "box":
            if len(params) != 3:
                raise Exception("params must be length 3 for box geometry (input [w, h, l])")
            self.u1 = params[0]*(self.u1 - 0.5)
            self.u2 = params[1]*(self.u2 - 0.5)
            self.u3 = params[2]*(self.u3 - 0.5)
        elif geometry == "sphere":
            if len(params) != 1:
                raise Exception("params must be length 1 for sphere geometry (input [radius])")
            self.u1 = params[0]*self.u1
            self.u2 = np.arccos(1 - 2*self.u2)
            self.u3 = 2*np.pi*self.u3
        elif geometry == "cylinder":
            if len(params) != 2:
                raise Exception("params must be length 2 for cylinder geometry (input [radius, height])")
            self.u1 = params[0]*self.u1
            self.u2 = 2*np.pi*self.u3
            self.u3 = params[1]*(self.u3 - 0.5)
        else:
            raise Exception("Geometry %s not found", geometry)

    def l_cart(self, x, y, z):
        return np.sqrt(x**2 + y**2 + (z - self.l0)**2)

    def l_cyl(self, r, theta, z):
        return np.sqrt(r**2 + self.l0**2 - 2*r*self.l0*np.cos(theta) + z**2)

    def l_sph(self, r, theta, phi):
        return np.sqrt(r**2 + self.l0**2 - 2*r*self.l0*np.sin(theta)*np.cos(phi))

    def point_table(self):
        """(distances l_k, jacobian factors J_k, volume factor V): integrate(f) = V * sum_k J_k f(l_k) / n_samples
        (materials.py:98-124)."""
        if self.geom == "box":
            return (self.l_cart(self.u1, self.u2, self.u3), np.ones(self.n_samples),
                    self.params[0]*self.params[1]*self.params[2])
        if self.geom == "sphere":
            return (self.l_sph(self.u1, self.u2, self.u3), self.u1**2 * np.sin(self.u2), 4*np.pi*self.params[0]**3 / 3)
        return (self.l_cyl(self.u1, self.u2, self.u3), self.u1, self.params[1]*np.pi*self.params[0]**2)

    def integrate(self, f_int):
        """Host evaluation for an arbitrary callable (scripts); the flux classes integrate on the device."""
        ls, jac, vol = self.point_table()
        return vol * np.sum(jac * f_int(ls)) / self.n_samples

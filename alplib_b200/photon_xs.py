"""Tabulated photon cross sections (NIST XCOM).  The table is loaded and de-duplicated on the host like the
reference (photon_xs.py:18-43); the log-log interpolation runs on the device (alpk_abs_sigma_mev and the per-row
kernels, which keep the table in shared memory)."""
import numpy as np
import torch

from . import _native as N
from .constants import AVOGADRO, MEV2_CM2
from .materials import Material, data_path
from .runtime import Runtime

_XS_DIM = {"NaI": 149.89 / AVOGADRO, "CsI": 259.81 / AVOGADRO, "CH2": 14.027 / AVOGADRO,
           "N2": 28.0 / AVOGADRO, "O2": 32.0 / AVOGADRO, "TeO2": 32.0 / AVOGADRO}      # photon_xs.py:27-38


_TABLE_CACHE = {}
_CLEAN_CACHE = {}
_DEVICE_TABLES = {}


def _load_table(path, skip_header):
    """np.genfromtxt(path, skip_header=...) with the parsed file kept in memory (the reference re-reads it per object)."""
    key = (path, skip_header)
    t = _TABLE_CACHE.get(key)
    if t is None:
        t = _TABLE_CACHE[key] = np.genfromtxt(path, skip_header=skip_header)
    return t.copy()


class AbsCrossSection:
    """Total photon absorption cross section vs energy [MeV] (photon_xs.py:18-52)."""

    def __init__(self, material: Material):
        self.xs_dim = _XS_DIM.get(material.mat_name, 1e-24)          # barns -> cm2, or cm2/g -> cm2/molecule
        self.mat_name = material.mat_name
        path = data_path("photon_absorption", "photon_abs_" + self.mat_name + ".txt")
        cached = _CLEAN_CACHE.get((path, self.xs_dim))
        if cached is None:
            self.pe_data = _load_table(path, 3)
            self.cleanPEData()
            _CLEAN_CACHE[(path, self.xs_dim)] = (self.pe_data.copy(), self.log10_e.copy(), self.log10_xs.copy())
        else:                                    # scans build one flux object per grid point: skip unique/log10
            self.pe_data, self.log10_e, self.log10_xs = (a.copy() for a in cached)
        # device copies are shared by every instance built from the same table (scans build one flux per point)
        self._dev = _DEVICE_TABLES.setdefault((self.mat_name, self.xs_dim), {})

    def cleanPEData(self):
        # keep the first row of every duplicated edge energy (photon_xs.py:42-43)
        self.pe_data = self.pe_data[np.unique(self.pe_data[:, 0], return_index=True)[1]]
        self.log10_e = np.log10(self.pe_data[:, 0])
        self.log10_xs = np.log10(self.xs_dim * self.pe_data[:, 1])

    def device_table(self, rt: Runtime):
        """(log10 E, log10 sigma) on the device of `rt`, uploaded once."""
        t = self._dev.get(rt.index)
        if t is None:
            t = self._dev[rt.index] = (rt.upload(self.log10_e), rt.upload(self.log10_xs), int(self.log10_e.shape[0]))
        return t

    def _sigma_mev_device(self, E, rt=None):
        rt = rt or Runtime.get()
        e = np.atleast_1d(np.asarray(E, dtype=np.float64))
        le, lx, nt = self.device_table(rt)
        d_e = rt.upload(e.ravel())
        out = rt.empty(e.size)
        N.call("alpk_abs_sigma_mev", N.ptr(le), N.ptr(lx), nt, N.ptr(d_e), e.size, N.ptr(out), rt.stream())
        return out.cpu().numpy().reshape(e.shape)

    def sigma_mev(self, E):
        r = self._sigma_mev_device(E)
        return r if np.ndim(E) else np.float64(r[0])

    def sigma_cm2(self, E):
        """10**interp(log10 E, log10 E_i, log10(xs_dim*sigma_i), 0, 0) (photon_xs.py:45-46) -- its own lookup, not
        sigma_mev*MEV2_CM2 (a divide and a multiply would move the last bit)."""
        rt = Runtime.get()
        le, lx, nt = self.device_table(rt)
        return _table_sigma(rt, 0, le, lx, nt, E)

    def mu(self, E, n):
        return self.sigma_cm2(E) * n


def _table_sigma(rt, kind, le, lx, nt, E):
    e = np.atleast_1d(np.asarray(E, dtype=np.float64))
    with rt.activate():
        d_e = rt.upload(e.ravel())
        out = rt.empty(e.size)
        N.call("alpk_table_sigma", kind, N.ptr(le), N.ptr(lx), nt, N.ptr(d_e), e.size, N.ptr(out), rt.stream())
    r = rt.to_host(out).reshape(e.shape)
    return r if np.ndim(E) else np.float64(r[0])


class ALPPairProdutionCrossSection:
    """Tabulated ALP pair-production cross section a e- -> e- e+ e- on argon, rescaled by (Z/18)^2 (reference:
    photon_xs.py:95-125, name kept as spelled there).  The table is picked by the closest mass control point at or
    below `ma`, exactly as the reference picks it."""

    def __init__(self, material: Material, ma=0.1):
        self.xs_dim = 1e-24
        self.mat_name = material.mat_name
        mass_extension = ["100keV", "400keV", "500keV", "600keV", "700keV", "800keV", "900keV", "1MeV", "1.1MeV"]
        mass_control_points = [0.1, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1]
        mass_idx = np.arange(0, 9, 1)
        idx_closest = np.clip(int(np.interp(ma, mass_control_points, mass_idx)), a_min=0, a_max=9)
        self.argon_rescale = np.power(material.z[0]/18, 2)
        self.xs_data = _load_table(data_path("alp_pair_production", "pair_production_xs_table_" + mass_extension[idx_closest] + ".txt"), 0)
        self.log10_e = np.log10(self.xs_data[:, 0])
        self.log10_xs_mev = np.log10(self.argon_rescale * self.xs_dim * self.xs_data[:, 1] / MEV2_CM2)
        self.log10_xs_cm2 = np.log10(self.argon_rescale * self.xs_dim * self.xs_data[:, 1])
        self._dev = {}
        self._dev_cm2 = {}

    def device_table(self, rt: Runtime):
        t = self._dev.get(rt.index)
        if t is None:
            t = self._dev[rt.index] = (rt.upload(self.log10_e), rt.upload(self.log10_xs_mev), int(self.log10_e.shape[0]))
        return t

    def sigma_cm2(self, E):
        """heaviside(E-2me,0)*10**interp(log10 E; log10(argon_rescale*xs_dim*sigma_i), left=-inf) (photon_xs.py:117-120)."""
        rt = Runtime.get()
        t = self._dev_cm2.get(rt.index)
        if t is None:
            t = self._dev_cm2[rt.index] = (rt.upload(self.log10_e), rt.upload(self.log10_xs_cm2), int(self.log10_e.shape[0]))
        return _table_sigma(rt, 1, t[0], t[1], t[2], E)

    def sigma_mev(self, E):
        """The same in MeV^-2 (photon_xs.py:122-125): the lookup pair_production() evaluates per sample."""
        rt = Runtime.get()
        le, lx, nt = self.device_table(rt)
        return _table_sigma(rt, 1, le, lx, nt, E)

    def mu(self, E, n):  # atomic number density in cm^-3
        return self.sigma_cm2(E) * n

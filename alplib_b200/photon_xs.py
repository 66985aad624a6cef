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
        self.pe_data = _load_table(data_path("photon_absorption", "photon_abs_" + self.mat_name + ".txt"), 3)
        self.cleanPEData()
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
        return self.sigma_mev(E) * MEV2_CM2

    def mu(self, E, n):
        return self.sigma_cm2(E) * n

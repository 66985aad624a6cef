"""
Squared matrix elements of the 2 -> 2 processes on the hot path -- drop-in for the reference's `matrix_element.py`
classes MatrixElement2, M2Compton (:629-645), M2InverseCompton (:650-666), M2InversePrimakoff (:671-686),
M2Primakoff (:690-705) and M2AssociatedProduction (:709-734).

The objects are host-side descriptors (masses + model parameters); `__call__(s, t, coupling_product=1.0)` and
`Scatter2to2MC` evaluate them on the device (libalpk `alpk_m2_eval` / `alpk_scatter2to2`).  A user-defined subclass of
`MatrixElement2` with a Python `__call__` cannot run inside a CUDA kernel: `Scatter2to2MC` rejects it (there is no CPU
fallback).
"""
import ctypes as C

import numpy as np
from numpy import pi, sqrt

from . import _native as N
from .constants import ALPHA, M_E
from .form_factors import AtomicElasticFF, atomic_ff_a2
from .runtime import Runtime

__all__ = ["MatrixElement2", "M2Compton", "M2InverseCompton", "M2InversePrimakoff", "M2Primakoff",
           "M2AssociatedProduction"]

INVERSE_PRIMAKOFF, PRIMAKOFF, INVERSE_COMPTON, COMPTON, ASSOCIATED_PRODUCTION = range(5)


class MatrixElement2:
    """reference: matrix_element.py:11-19"""
    _MODEL = None

    def __init__(self, m1, m2, m3, m4):
        self.m1 = m1
        self.m2 = m2
        self.m3 = m3
        self.m4 = m4

    # -- device descriptor ---------------------------------------------------------------------------------------------
    def _m2_struct(self, s, coupling_product=1.0):
        """alpk_m2_t of this matrix element (s enters only the soft-photon cut of associated production)."""
        if self._MODEL is None:
            raise NotImplementedError(
                f"{type(self).__name__}: only the built-in matrix elements (M2Compton, M2InverseCompton, M2InversePrimakoff, "
                "M2Primakoff, M2AssociatedProduction) have a CUDA implementation; alplib_b200 has no CPU fallback")
        m = N.M2T()
        ma = getattr(self, "ma", 0.0)
        mN = getattr(self, "mN", 0.0)
        z = getattr(self, "z", 1)
        m.model = self._MODEL
        m.ma2, m.ma4 = float(ma**2), float(ma**4)
        m.mN2, m.mN4 = float(mN**2), float(mN**4)
        m.z = float(z)
        m.ff_a2 = atomic_ff_a2(z) if self._MODEL in (INVERSE_PRIMAKOFF, PRIMAKOFF) else 0.0
        m.me2, m.me4 = float(M_E**2), float(M_E**4)
        m.pref = float(self._prefactor(coupling_product))
        m.tmax = float(self._tmax(s)) if self._MODEL == ASSOCIATED_PRODUCTION else 0.0
        return m

    def _prefactor(self, coupling_product):
        raise NotImplementedError

    def _eval(self, s, t, coupling_product, den=0.0):
        sa, ta = np.broadcast_arrays(np.asarray(s, dtype=np.float64), np.asarray(t, dtype=np.float64))
        shape = ta.shape
        rt = Runtime.get(None)
        with rt.activate():
            d_t = rt.upload(np.ascontiguousarray(ta).ravel())
            d_s = rt.upload(np.ascontiguousarray(sa).ravel())
            out = rt.empty(ta.size)
            if self._MODEL == ASSOCIATED_PRODUCTION and np.ndim(s) != 0 and np.unique(sa).size > 1:
                raise NotImplementedError("M2AssociatedProduction: one value of s per call (its soft-photon cut depends on s)")
            m = self._m2_struct(float(sa.ravel()[0]) if sa.size else 0.0, coupling_product)
            N.call("alpk_m2_eval", C.byref(m), N.ptr(d_s), 1, N.ptr(d_t), int(ta.size), float(den), N.ptr(out), rt.stream())
        res = out.cpu().numpy().reshape(shape)
        return res if len(shape) else np.float64(res)

    def __call__(self, s, t, coupling_product=1.0):
        if self._MODEL is None:
            return 0.0                                  # the reference's base class (matrix_element.py:18-19)
        return self._eval(s, t, coupling_product)


class M2Compton(MatrixElement2):
    """Compton scattering (gamma + e- -> a + e-); reference: matrix_element.py:629-645"""
    _MODEL = COMPTON

    def __init__(self, ma, z):
        super().__init__(0, M_E, ma, M_E)
        self.ma = ma
        self.z = z

    def _prefactor(self, coupling_product):
        return self.z*4*pi*ALPHA * coupling_product**2


class M2InverseCompton(MatrixElement2):
    """Compton scattering (a + e- -> gamma + e-); reference: matrix_element.py:650-666"""
    _MODEL = INVERSE_COMPTON

    def __init__(self, ma, z):
        super().__init__(ma, M_E, 0, M_E)
        self.ma = ma
        self.z = z

    def _prefactor(self, coupling_product):
        return self.z*4*pi*ALPHA * coupling_product**2


class M2InversePrimakoff(MatrixElement2):
    """Inverse Primakoff scattering (a + N -> gamma + N); reference: matrix_element.py:671-686"""
    _MODEL = INVERSE_PRIMAKOFF

    def __init__(self, ma, mN, z):
        super().__init__(ma, mN, 0, mN)
        self.mN = mN
        self.ma = ma
        self.z = z
        self.ff2 = AtomicElasticFF(z)

    def _prefactor(self, coupling_product):
        return 2 * coupling_product**2


class M2Primakoff(MatrixElement2):
    """Primakoff scattering (gamma + N -> a + N); reference: matrix_element.py:690-705"""
    _MODEL = PRIMAKOFF

    def __init__(self, ma, mN, z):
        super().__init__(0, mN, ma, mN)
        self.mN = mN
        self.ma = ma
        self.z = z
        self.ff2 = AtomicElasticFF(z)

    def _prefactor(self, coupling_product):
        return coupling_product**2


class M2AssociatedProduction(MatrixElement2):
    """Associated ALP production (e+ + e- -> a + gamma); reference: matrix_element.py:709-734"""
    _MODEL = ASSOCIATED_PRODUCTION

    def __init__(self, ma):
        super().__init__(M_E, M_E, ma, 0.0)
        self.ma = ma

    def _prefactor(self, coupling_product):
        return (pi*ALPHA) * coupling_product**2

    @staticmethod
    def _tmax(s):
        # Phase space cutoff (matrix_element.py:728-731)
        with np.errstate(all="ignore"):
            p_pos_cm = sqrt((np.power(s - 2*M_E**2, 2) - 4*np.power(M_E, 4))/(4*s))
            e_pos_cm = sqrt(p_pos_cm**2 + M_E**2)
            return M_E**2 - 2*0.01*e_pos_cm*(e_pos_cm)  # soft-photon cutoff at 1%

"""Form factors used on the hot path (reference: form_factors.py).  Only the atomic elastic form factor is needed by the
2 -> 2 matrix elements of `matrix_element.py`; it is evaluated on the device (libalpk `alpk_atomic_ff2`)."""
import numpy as np

from . import _native as N
from .constants import M_E
from .runtime import Runtime


def atomic_ff_a2(z):
    """a**2 with a = 184.15*2.718**(-1/2)*z**(-1/3)/M_E (form_factors.py:63), as the reference evaluates it."""
    a = 184.15*np.power(2.718, -1/2)*np.power(z, -1/3) / M_E
    return float(a**2)


class AtomicElasticFF:
    """square of the atomic elastic form factor (reference: form_factors.py:54-64)"""
    def __init__(self, z):
        self.z = z

    def __call__(self, q):
        qa = np.ascontiguousarray(q, dtype=np.float64)
        rt = Runtime.get(None)
        with rt.activate():
            d_q = rt.upload(qa.ravel())
            out = rt.empty(qa.size)
            N.call("alpk_atomic_ff2", N.ptr(d_q), int(qa.size), float(self.z), atomic_ff_a2(self.z), N.ptr(out), rt.stream())
        res = out.cpu().numpy().reshape(qa.shape)
        return res if qa.ndim else np.float64(res)

"""Host-side evaluation of the per-call scalar constants handed to the kernels.

Each scalar is computed with the same Python/NumPy expression as the reference (cited file:line) so that it is
bit-identical to the value the reference would use; the kernels then only do the per-row / per-sample arithmetic.
"""
import numpy as np

from . import _native as N
from .constants import ALPHA, AVOGADRO, C_LIGHT, HBARC, M_E, METER_BY_MEV, S_PER_DAY, pi

R0_SCREEN = 2.2e-10 / METER_BY_MEV      # prod_xs.py:99 / det_xs.py:37 default screening parameter


def prop_params(ma, width, rescale=1.0, det_dist=4.0, det_length=0.2, geom_accept=None, timing_window=None, fast=False):
    """fluxes.py:43-65 scalars.  geom_accept None <=> is_isotropic False (fluxes.py:159)."""
    p = N.PropT()
    p.ma = float(ma)
    p.ma2 = float(ma ** 2)
    p.width = float(width)
    p.rescale = float(rescale)
    p.neg_dist_mbm = float(-det_dist / METER_BY_MEV)
    p.neg_len_mbm = float(-det_length / METER_BY_MEV)
    if geom_accept is not None:
        p.apply_geom = 1
        p.geom = float(geom_accept)
        p.geom32 = float(np.float32(geom_accept))
        # NEP 50: a Python float factor multiplies the float32 weights in float32; an np.float64 (or other
        # non-weak) factor is applied in float64 and the result cast back (in-place *= on a float32 array).
        p.geom_f64 = 1 if isinstance(geom_accept, np.floating) and not isinstance(geom_accept, np.float32) else 0
    else:
        p.apply_geom = 0
        p.geom = 1.0
        p.geom32 = 1.0
        p.geom_f64 = 0
    if timing_window is not None:
        p.use_tof = 1
        p.tof_light = float(det_dist / (1e-2 * C_LIGHT))
        p.timing_window = float(timing_window)
    else:
        p.use_tof = 0
        p.tof_light = 0.0
        p.timing_window = 0.0
    p.det_dist = float(det_dist)
    # fast mode: L/(hbar c v tau) = (L/hbar c) * ma*Gamma / p_a  (exponents = constant / p_a)
    p.fast = 2 if fast == "fp32" else (1 if fast else 0)
    p.k1 = float(p.neg_dist_mbm * ma * width) if width > 0.0 else 0.0
    p.k2 = float(p.neg_len_mbm * ma * width) if width > 0.0 else 0.0
    return p


def geom_acceptance(det_area, det_dist):
    return det_area / (4 * pi * det_dist ** 2)          # fluxes.py:160


def event_params(days_exposure, threshold):
    """generators.py:88-92: days_exposure * S_PER_DAY * w32 * heaviside(E - threshold, 1)."""
    e = N.EventT()
    ds = days_exposure * S_PER_DAY
    e.days_sec = float(ds)
    e.days_sec32 = float(np.float32(ds))
    e.days_f64 = 1 if isinstance(ds, np.floating) and not isinstance(ds, np.float32) else 0
    e.threshold = float(threshold)
    return e


def compton_params(ma, g, z, n_samples, fast=False, rng_rounds=10):
    c = N.ComptonT()
    c.ma = float(ma)
    c.ma2 = float(ma ** 2)
    c.aa = float(g ** 2 / 4 / pi)                        # prod_xs.py:157
    c.z = float(z)
    c.n_samples = float(n_samples)
    c.fast = 1 if fast else 0
    c.rng_rounds = int(rng_rounds)
    return c


def primakoff_consts(g, z, r0=R0_SCREEN):
    """prod_xs.py:102 / det_xs.py:40: prefactor = (g*z)**2/(2*137), r0**2."""
    return float((g * z) ** 2 / (2 * 137)), float(r0 ** 2)


def icompton_consts(ma, g, z):
    """det_xs.py:89-98 scalars: {ma, ma**2, ma**4, e_thr, z*ALPHA*(g/M_E)**2, ma**4 + 2 (ma M_E)**2}."""
    ethr = ma * np.sqrt(2 * M_E ** 2 + ma ** 2) / (2 * M_E)
    zfac = z * ALPHA * np.power(g / M_E, 2)
    c2 = ma ** 4 + 2 * np.power(ma * M_E, 2)
    arr = (N.C.c_double * 6)(float(ma), float(ma ** 2), float(ma ** 4), float(ethr), float(zfac), float(c2))
    return arr


def iprimakoff_consts(ma, g, z, r0=R0_SCREEN):
    pref, r02 = primakoff_consts(g, z, r0)
    return (N.C.c_double * 6)(float(ma), float(ma ** 2), pref, r02, 0.0, 0.0)


_GL = None


def gauss_legendre():
    """Nodes/weights of the 32-point Gauss-Legendre rule used for the track-length integral (fmath.py:33-36).  Computed
    once: numpy's leggauss (companion-matrix eigenvalues + Newton step) costs 0.4 ms, more than the kernels it feeds."""
    global _GL
    if _GL is None:
        x, w = np.polynomial.legendre.leggauss(N.GL_POINTS)
        x, w = np.ascontiguousarray(x), np.ascontiguousarray(w)
        x.setflags(write=False); w.setflags(write=False)
        _GL = (x, w)
    return _GL


def brem_params(ma, g, z, rad_length, n_samples, boson_type="pseudoscalar", max_t=5.0, use_track_length=True, fast=False,
                rng_rounds=10):
    """fluxes.py:276-337 + prod_xs.py:209-221, 313-323 scalars (z is the np.int64 of Material.z[0])."""
    from numpy import log, power
    b = N.BremT()
    b.ma = float(ma)
    b.ma2 = float(ma ** 2)
    b.log10_ma = float(np.log10(ma))
    b.ep_min = float(max(ma, M_E))                                   # fluxes.py:346
    ntad = rad_length * AVOGADRO / (2 * z)                            # fluxes.py:290
    b.nt = float(ntad * HBARC ** 2)                                   # :330
    r0 = ALPHA / M_E
    b.pref_num = float(2 * r0 ** 2 * g ** 2 / 4 / pi)                 # prod_xs.py:217 (before / Ee)
    ln_el = log(184 * power(z, -1 / 3))
    ln_inel = log(1194 * power(z, -2 / 3))
    b.zl = float(z ** 2 * ln_el + z * ln_inel)
    b.zz = float(z ** 2 + z)
    chi = z ** 2 * ln_el + z * ln_inel
    b.pref_vec = float(chi * (4 * ALPHA ** 2 * g ** 2) / (4 * pi))    # prod_xs.py:322
    b.ln10 = float(np.log(10))
    b.n_samples = float(n_samples)
    b.t_arg = float(4 * max_t / 3)                                    # fluxes.py:266
    if boson_type not in ("pseudoscalar", "vector"):
        # the reference leaves diff_br unbound for any other string (fluxes.py:328-337)
        raise UnboundLocalError("cannot access local variable 'diff_br' where it is not associated with a value")
    b.vector = 1 if boson_type == "vector" else 0
    b.use_track_length = 1 if use_track_length else 0
    b.fast = 1 if fast else 0
    b.rng_rounds = int(rng_rounds)
    return b


def pairann_params(ma, ge, z, rad_length, n_samples, fast=False, rng_rounds=10):
    """fluxes.py:484,518 + matrix_element.py:717-734 scalars."""
    p = N.PairAnnT()
    p.ma = float(ma)
    p.ma2 = float(ma ** 2)
    p.me2 = float(M_E ** 2)
    p.me4 = float(M_E ** 4)
    p.pref = float((pi * ALPHA) * 1.0 ** 2)
    p.wconst = 0.0
    p.z = float(z)
    p.ge2 = float(ge ** 2)
    p.ntad = float(rad_length * AVOGADRO / (2 * z))
    p.hbarc2 = float(HBARC ** 2)
    p.n_samples = float(n_samples)
    p.fast = 1 if fast else 0
    p.rng_rounds = int(rng_rounds)
    return p

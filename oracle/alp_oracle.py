"""
CPU oracle for the alplib Monte-Carlo flux-and-event hot path.

TEST INFRASTRUCTURE ONLY.  This module is a NumPy/SciPy restatement of the reference algorithm
(athompson-git/alplib).  It may be imported ONLY by `tests/`, by `__graft_entry__.smoke()` and by the
`cpu_baseline` / `--impl reference` legs of `bench.py` -- always as the checker or as the timed CPU
baseline, never as part of the product path (`alplib_b200/` never imports it).

Parity status: PINNED.  The reference publishes no golden vectors for this path, so the
oracle is pinned against outputs of the unmodified reference itself, run in the build container with
identical replayed uniforms (`tests/golden/make_golden.py` -> `tests/golden/*.npz`) and against the
known-answer values of SURVEY.md section 8c (`tests/test_oracle_golden.py`).

Every function cites the reference file:line (relative to the reference repository root) it follows.
Operation order follows the reference expression by expression so that FP64 results agree to the last few
ulp (the remaining differences are NumPy scalar-vs-array libm paths, SURVEY.md section 7 hard part 6).

Third-party arithmetic used by the reference on this path and reused here: NumPy (`np.interp`, `np.unique`,
ufuncs), SciPy `integrate.quad` (fmath.py:36) and `special.gamma` (fluxes.py:259); reference pins
numpy==2.3.3 / scipy==1.16.2 (requirements.txt:2-3).
"""

import json
import os

import numpy as np
from numpy import pi
from scipy.integrate import quad
from scipy.special import gamma as _gamma

# ---------------------------------------------------------------------------------------------------------
# constants.py:14-66 (values restated bit for bit; KAT in tests/test_oracle_golden.py)
# ---------------------------------------------------------------------------------------------------------
ALPHA = 1 / 137                       # constants.py:14
C_LIGHT = 2.99792458e10               # constants.py:20
HBARC = 1.97236e-11                   # constants.py:22
AVOGADRO = 6.022e23                   # constants.py:29
M_E = 0.511                           # constants.py:32
METER_BY_MEV = 6.58212e-22 * 2.998e8  # constants.py:60
MEV2_CM2 = (METER_BY_MEV * 100) ** 2  # constants.py:65
S_PER_DAY = 3600 * 24                 # constants.py:66
MEV_PER_KG = 5.6095887e29             # constants.py:61

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "alplib_b200", "data")


class UniformSource:
    """Source of U[0,1) variates consumed in the reference's draw order.

    Wraps either a legacy `np.random.RandomState` (so `RandomState(s)` replays `np.random.seed(s)` of the
    reference bit for bit: legacy `uniform(lo,hi,n) = lo + (hi-lo)*random_sample(n)`) or a preset 1-D array.
    Everything handed out is logged so it can be injected into the CUDA kernels.
    """

    def __init__(self, src):
        self._rs = src if isinstance(src, np.random.RandomState) else None
        self._arr = None if self._rs is not None else np.asarray(src, dtype=np.float64).ravel()
        self._pos = 0
        self.log = []

    def take(self, n=None):
        k = 1 if n is None else int(n)
        if self._rs is not None:
            u = self._rs.random_sample(k)
        else:
            u = self._arr[self._pos:self._pos + k]
            if u.shape[0] != k:
                raise ValueError("UniformSource exhausted")
            self._pos += k
        self.log.append(u)
        return u[0] if n is None else u

    def uniform(self, lo, hi, n=None):
        # numpy legacy: loc + scale * double  (mtrand legacy_uniform)
        return lo + (hi - lo) * self.take(n)


# ---------------------------------------------------------------------------------------------------------
# materials.py:12-46  (host-side input container)
# ---------------------------------------------------------------------------------------------------------
class Material:
    def __init__(self, material_name, fiducial_mass=1.0, volume=1.0, density=1.0):
        with open(os.path.join(_DATA_DIR, "mat_params.json")) as f:
            mat_file = json.load(f)
        if material_name not in mat_file:
            raise Exception("No such detector in mat_params.json.")          # materials.py:46
        info = mat_file[material_name]
        self.mat_name = material_name
        self.iso = info["iso"]
        self.z = np.array(info["z"])
        self.n = np.array(info["n"])
        self.m = np.array(info["m"])
        self.frac = np.array(info["frac"])
        self.density = info["density"]
        self.rad_length = info["rad_length"]
        self.fid_mass = fiducial_mass
        self.volume = volume
        # materials.py:42-43 -- uses the ctor's density argument, not the JSON one
        self.ntargets = density * volume / (np.dot(self.m, self.iso * self.frac) / MEV_PER_KG / 1e-3)
        self.ndensity = density / (np.dot(self.m, self.iso * self.frac) / MEV_PER_KG / 1e-3)


# ---------------------------------------------------------------------------------------------------------
# photon_xs.py:18-52  AbsCrossSection
# ---------------------------------------------------------------------------------------------------------
_XS_DIM = {"NaI": 149.89 / AVOGADRO, "CsI": 259.81 / AVOGADRO, "CH2": 14.027 / AVOGADRO,
           "N2": 28.0 / AVOGADRO, "O2": 32.0 / AVOGADRO, "TeO2": 32.0 / AVOGADRO}   # photon_xs.py:27-38


class AbsCrossSection:
    def __init__(self, material):
        name = material.mat_name if hasattr(material, "mat_name") else str(material)
        self.mat_name = name
        self.xs_dim = _XS_DIM.get(name, 1e-24)                                       # photon_xs.py:20
        fpath = os.path.join(_DATA_DIR, "photon_absorption", "photon_abs_" + name + ".txt")
        pe = np.genfromtxt(fpath, skip_header=3)                                     # photon_xs.py:25
        self.pe_data = pe[np.unique(pe[:, 0], return_index=True)[1]]                 # photon_xs.py:42-43
        self.log10_e = np.log10(self.pe_data[:, 0])
        self.log10_xs = np.log10(self.xs_dim * self.pe_data[:, 1])

    def sigma_mev(self, E):                                                          # photon_xs.py:48-49
        return 10 ** np.interp(np.log10(E), self.log10_e, self.log10_xs, left=0.0, right=0.0) / MEV2_CM2


# ---------------------------------------------------------------------------------------------------------
# decay.py:10-52  widths
# ---------------------------------------------------------------------------------------------------------
def W_gg(g_agamma, ma):                                                              # decay.py:10-13
    return g_agamma ** 2 * ma ** 3 / (64 * pi)


def W_ee(g_ae, ma):                                                                  # decay.py:26-29
    return g_ae ** 2 * ma * np.sqrt(1 - (2 * M_E / ma) ** 2) / (8 * pi) \
        if 1 - 4 * (M_E / ma) ** 2 > 0 else 0.0


def fp2(tau):                                                                        # decay.py:35-37
    return np.arcsin(1 / np.sqrt(tau)) ** 2 if tau >= 1 \
        else pi ** 2 / 4 + np.log((1 + np.sqrt(1 - tau)) / (1 - np.sqrt(1 - tau))) ** 2 / 4


def b1(tau):                                                                         # decay.py:42-43
    return 1 - tau * fp2(tau)


def W_gg_loop(g_af, ma, mf):                                                         # decay.py:48-52
    g_agamma_eff = g_af * ALPHA * b1(np.power(2 * mf / ma, 2)) / (4 * pi)
    return W_gg(g_agamma_eff, ma)


# ---------------------------------------------------------------------------------------------------------
# prod_xs.py  production cross sections
# ---------------------------------------------------------------------------------------------------------
def primakoff_sigma(eg, g, ma, z, r0=2.2e-10 / METER_BY_MEV):                        # prod_xs.py:99-104
    prefactor = (g * z) ** 2 / (2 * 137)
    eta2 = r0 ** 2 * eg ** 2
    return np.heaviside(eg - ma, 0.0) * prefactor * (((2 * eta2 + 1) / (4 * eta2)) * np.log(1 + 4 * eta2) - 1)


def compton_dsigma_dea(ea, eg, g, ma, z=1):                                          # prod_xs.py:155-169
    a = 1 / 137
    aa = g ** 2 / 4 / pi
    s = 2 * M_E * eg + M_E ** 2
    x = ((ma ** 2 / (2 * eg * M_E)) - ea / eg + 1)
    xmin = ((s - M_E ** 2) * (s - M_E ** 2 + ma ** 2)
            - (s - M_E ** 2) * np.sqrt((s - M_E ** 2 + ma ** 2) ** 2 - 4 * s * ma ** 2)) / (2 * s * (s - M_E ** 2))
    xmax = ((s - M_E ** 2) * (s - M_E ** 2 + ma ** 2)
            + (s - M_E ** 2) * np.sqrt((s - M_E ** 2 + ma ** 2) ** 2 - 4 * s * ma ** 2)) / (2 * s * (s - M_E ** 2))
    thresh = np.heaviside(eg - ma, 0.0) * np.heaviside(x - xmin, 0.0) * np.heaviside(xmax - x, 0.0)
    return z * thresh * (1 / eg) * pi * a * aa / (s - M_E ** 2) * (
        x / (1 - x) * (-2 * ma ** 2 / (s - M_E ** 2) ** 2 * (s - M_E ** 2 / (1 - x) - ma ** 2 / x) + x))


def brem_dsigma_dea(Ea, Ee, g, ma, z):                                               # prod_xs.py:209-221
    r0 = ALPHA / M_E
    x = Ea / Ee
    f = np.power(ma / (x * M_E), 2) * (1 - x)
    ln_el = np.log(184 * np.power(z, -1 / 3))
    ln_inel = np.log(1194 * np.power(z, -2 / 3))
    prefactor = 2 * r0 ** 2 * g ** 2 / 4 / pi / Ee
    phase_space = ((x * (1 + f / 1.5) / np.power(1 + f, 2)) * (z ** 2 * ln_el + z * ln_inel)
                   + x * (z ** 2 + z) * ((1 + f) * np.log(1 + f) / (3 * f ** 2)
                                         - (1 + 4 * f + 2 * f ** 2) / (3 * f * np.power(1 + f, 2))))
    return prefactor * phase_space * np.heaviside(phase_space, 0.0)


def brem_dsigma_dx_vector(x, coupling, ma, z):                                       # prod_xs.py:313-323
    ln_el = np.log(184 * np.power(z, -1 / 3))
    ln_inel = np.log(1194 * np.power(z, -2 / 3))
    chi = z ** 2 * ln_el + z * ln_inel
    prefactor = chi * (4 * ALPHA ** 2 * coupling ** 2) / (4 * pi)
    return prefactor * (1 - x + x ** 2 / 3) / ((ma ** 2 * (1 - x) / x) + x * M_E ** 2)


# ---------------------------------------------------------------------------------------------------------
# det_xs.py  detection cross sections
# ---------------------------------------------------------------------------------------------------------
def iprimakoff_sigma(ea, g, ma, z, r0=2.2e-10 / METER_BY_MEV):                       # det_xs.py:37-42
    prefactor = (g * z) ** 2 / (2 * 137)
    eta2 = r0 ** 2 * (ea ** 2 - ma ** 2)
    with np.errstate(all="ignore"):
        return np.heaviside(ea - ma, 0.0) * prefactor * (((2 * eta2 + 1) / (4 * eta2)) * np.log(1 + 4 * eta2) - 1)


def icompton_sigma(ea, ma, g, z=1):                                                  # det_xs.py:89-98
    with np.errstate(all="ignore"):
        y = 2 * M_E * ea + ma ** 2
        pa = np.sqrt(np.clip(ea ** 2 - ma ** 2, a_min=0.0, a_max=np.inf))
        prefactor = np.heaviside(ea - ma, 0.0) \
            * np.heaviside(ea - ma * np.sqrt(2 * M_E ** 2 + ma ** 2) / (2 * M_E), 0.0) \
            * z * ALPHA * np.power(g / M_E, 2) / (8 * pa)
        return np.nan_to_num(prefactor * ((2 * M_E ** 2 * (M_E + ea) * y) / np.power(M_E ** 2 + y, 2)
                                          + (4 * M_E * (ma ** 4 + 2 * np.power(ma * M_E, 2) - np.power(2 * M_E * ea, 2)))
                                          / (y * (M_E ** 2 + y))
                                          + np.log((M_E + ea + pa) / (M_E + ea - pa))
                                          * (np.power(2 * M_E * pa, 2) + ma ** 4) / (pa * y)))


# ---------------------------------------------------------------------------------------------------------
# fmath.py:33-36, fluxes.py:257-266  track-length helpers
# ---------------------------------------------------------------------------------------------------------
def nu_integral(x, alpha, T_max=np.inf):                                             # fmath.py:33-36
    f = lambda t: np.power(x, t + alpha) / _gamma(t + 1 + alpha)
    return quad(f, 0.0, T_max)[0]


def track_length_prob(Ei, Ef, t):                                                    # fluxes.py:257-259
    b = 4 / 3
    with np.errstate(all="ignore"):
        return np.heaviside(Ei - Ef, 1.0) * abs(np.power(np.log(Ei / Ef), b * t - 1) / (Ei * _gamma(b * t)))


def track_length_integrated_prob(Ei, Ef, T_max=5.0):                                 # fluxes.py:264-266
    return (3 / 4) * np.heaviside(Ei - Ef, 1.0) * nu_integral(np.log(Ei / Ef), -1, 4 * T_max / 3) / Ei


# ---------------------------------------------------------------------------------------------------------
# fluxes.py:43-65 + 153-162  propagate (+ isotropic acceptance)
# ---------------------------------------------------------------------------------------------------------
def propagate(e_a, wgt, ma, decay_width, rescale_factor, det_dist, det_length,
              timing_window=None, geom_accept=None, geom_accept_f64=False):
    """Returns dict with the FP64 pre-cast weights (`decay_w64`, `scatter_w64`, before the acceptance factor)
    and the reference's float32 attributes (`decay_w`, `scatter_w`, after fluxes.py:159-162).

    `geom_accept` None == `is_isotropic=False`.  The in-place `*=` of fluxes.py:161-162 runs in float32 when
    `geom_accept` is a Python float (NEP 50 weak scalar) and in float64-then-cast when it is an np.float64
    (`geom_accept_f64=True`)."""
    e_a = np.asarray(e_a, dtype=np.float64)
    wgt = np.array(wgt, dtype=np.float64)
    with np.errstate(all="ignore"):
        p_a = np.sqrt(e_a ** 2 - ma ** 2)                                            # fluxes.py:48
        v_a = p_a / e_a                                                              # :49
        boost = e_a / ma                                                             # :50
        tau = boost / decay_width if decay_width > 0.0 else np.inf * np.ones_like(boost)   # :51
        if timing_window is not None:                                                # :54-58
            tof = det_dist / (v_a * 1e-2 * C_LIGHT)
            delta_tof = abs((det_dist / (1e-2 * C_LIGHT)) - tof)
            wgt = wgt * (delta_tof < timing_window)
        surv_prob = np.exp(-det_dist / METER_BY_MEV / v_a / tau)                     # :61
        decay_prob = (1 - np.exp(-det_length / METER_BY_MEV / v_a / tau))            # :62
        d64 = rescale_factor * wgt * surv_prob * decay_prob                          # :64
        s64 = rescale_factor * wgt * surv_prob                                       # :65
    d32 = np.asarray(d64, dtype=np.float32)
    s32 = np.asarray(s64, dtype=np.float32)
    if geom_accept is not None:                                                      # fluxes.py:159-162
        if geom_accept_f64:
            d32 = (d32.astype(np.float64) * np.float64(geom_accept)).astype(np.float32)
            s32 = (s32.astype(np.float64) * np.float64(geom_accept)).astype(np.float32)
        else:
            d32 = d32 * np.float32(geom_accept)
            s32 = s32 * np.float32(geom_accept)
    return {"decay_w64": d64, "scatter_w64": s64, "decay_w": d32, "scatter_w": s32}


# ---------------------------------------------------------------------------------------------------------
# materials.py:51-124 DetectorGeometry + fluxes.py:67-89 propagate_iso_vol_int
# ---------------------------------------------------------------------------------------------------------
class DetectorGeometry:
    def __init__(self, distance_to_source, geometry, params, n_samples, usrc):
        """usrc: UniformSource -- three blocks of n_samples draws (materials.py:58-60)."""
        self.l0, self.n_samples, self.geom, self.params = distance_to_source, n_samples, geometry, params
        u1, u2, u3 = usrc.uniform(0, 1, n_samples), usrc.uniform(0, 1, n_samples), usrc.uniform(0, 1, n_samples)
        if geometry == "box":                                                        # :64-70
            self.u1, self.u2, self.u3 = params[0] * (u1 - 0.5), params[1] * (u2 - 0.5), params[2] * (u3 - 0.5)
        elif geometry == "sphere":                                                   # :71-77
            self.u1, self.u2, self.u3 = params[0] * u1, np.arccos(1 - 2 * u2), 2 * pi * u3
        elif geometry == "cylinder":                                                 # :78-84 (azimuth from u3: reference quirk)
            self.u1, self.u2, self.u3 = params[0] * u1, 2 * pi * u3, params[1] * (u3 - 0.5)
        else:
            raise Exception("Geometry %s not found", geometry)

    def integrate(self, f_int):                                                      # :98-124
        if self.geom == "box":
            ls = np.sqrt(self.u1 ** 2 + self.u2 ** 2 + (self.u3 - self.l0) ** 2)
            return (self.params[0] * self.params[1] * self.params[2]) * np.sum(f_int(ls)) / self.n_samples
        if self.geom == "sphere":
            ls = np.sqrt(self.u1 ** 2 + self.l0 ** 2 - 2 * self.u1 * self.l0 * np.sin(self.u2) * np.cos(self.u3))
            return (4 * pi * self.params[0] ** 3 / 3) * np.sum(self.u1 ** 2 * np.sin(self.u2) * f_int(ls)) / self.n_samples
        ls = np.sqrt(self.u1 ** 2 + self.l0 ** 2 - 2 * self.u1 * self.l0 * np.cos(self.u2) + self.u3 ** 2)
        return (self.params[1] * pi * self.params[0] ** 2) * np.sum(self.u1 * f_int(ls)) / self.n_samples


def propagate_iso_vol_int(e_a, wgt, ma, decay_width, rescale_factor, det_dist, geom):
    """fluxes.py:67-89.  Returns f64 pre-cast weights and the float32 attributes."""
    e_a = np.asarray(e_a, dtype=np.float64)
    wgt = np.asarray(wgt, dtype=np.float64)
    with np.errstate(all="ignore"):
        p_a = np.sqrt(e_a ** 2 - ma ** 2)
        v_a = p_a / e_a
        boost = e_a / ma
        tau = boost / decay_width if decay_width > 0.0 else np.inf * np.ones_like(boost)
        surv = np.array([np.exp(-det_dist / METER_BY_MEV / v_a[i] / tau[i]) for i in range(len(v_a))])
        vol = []
        for i in range(len(e_a)):
            f = lambda x: (1 / METER_BY_MEV / v_a[i] / tau[i]) * np.exp(-x / METER_BY_MEV / v_a[i] / tau[i]) / (4 * pi * x ** 2)
            vol.append(geom.integrate(f))
        d64 = rescale_factor * wgt * np.array(vol)
        s64 = rescale_factor * wgt * surv
    return {"decay_w64": d64, "scatter_w64": s64, "decay_w": np.asarray(d64, dtype=np.float32),
            "scatter_w": np.asarray(s64, dtype=np.float32)}


def geom_acceptance(det_area, det_dist):                                             # fluxes.py:160
    return det_area / (4 * pi * det_dist ** 2)


# ---------------------------------------------------------------------------------------------------------
# fluxes.py:133-151  FluxPrimakoffIsotropic.simulate
# ---------------------------------------------------------------------------------------------------------
def primakoff_simulate(photon_flux, ma, g, target_z, abs_xs):
    photon_flux = np.asarray(photon_flux, dtype=np.float64)
    energy, flux = [], []
    for el in photon_flux:                                                           # fluxes.py:150-151
        eg, wgt = el[0], el[1]
        if eg < ma:                                                                  # :136
            continue
        xs = primakoff_sigma(eg, g, ma, target_z)                                    # :139
        br = xs / abs_xs.sigma_mev(eg)                                               # :140
        energy.append(eg)
        flux.append(wgt * br)                                                        # :142
    return np.array(energy, dtype=np.float64), np.array(flux, dtype=np.float64)


# ---------------------------------------------------------------------------------------------------------
# fluxes.py:210-233  FluxComptonIsotropic.simulate
# ---------------------------------------------------------------------------------------------------------
def compton_row_valid(photon_flux, ma):                                              # fluxes.py:214-216
    eg = np.asarray(photon_flux, dtype=np.float64)[:, 0]
    s = 2 * M_E * eg + M_E ** 2
    return ~(s < (M_E + ma) ** 2)


def compton_simulate(photon_flux, ma, g, target_z, abs_xs, n_samples, usrc):
    """Draw order: per row passing the s-threshold, `n_samples` uniforms (fluxes.py:218)."""
    photon_flux = np.asarray(photon_flux, dtype=np.float64)
    energy, flux = [], []
    with np.errstate(all="ignore"):
        for el in photon_flux:
            eg, wgt = el[0], el[1]
            s = 2 * M_E * eg + M_E ** 2                                              # :214
            if s < (M_E + ma) ** 2:
                continue
            ea = usrc.uniform(ma, eg, n_samples)                                     # :218
            mc_xs = (eg - ma) * compton_dsigma_dea(ea, eg, g, ma, target_z) / n_samples   # :219
            diff_br = mc_xs / abs_xs.sigma_mev(eg)                                   # :220
            energy.append(ea)
            flux.append(wgt * diff_br)                                               # :224
    if not energy:
        return np.zeros(0), np.zeros(0)
    return np.concatenate(energy), np.concatenate(flux)


# ---------------------------------------------------------------------------------------------------------
# fluxes.py:276-356  FluxBremIsotropic.simulate
# ---------------------------------------------------------------------------------------------------------
def brem_simulate(electron_flux, positron_flux, ma, g, target, n_samples, usrc, boson_type="pseudoscalar",
                  max_t=5.0, use_track_length=True):
    """Draw order (use_track_length): per e- row with E0 >= max(ma, M_E): 1 scalar (fluxes.py:350), then --
    only if ea_max > ma -- `n_samples` (fluxes.py:326).  Returns (E, flux, info) where info carries the
    per-row scalar draws and validity for uniform injection."""
    electron_flux = np.asarray(electron_flux, dtype=np.float64)
    positron_flux = np.asarray(positron_flux, dtype=np.float64)
    z = target.z[0]
    ntarget_area_density = target.rad_length * AVOGADRO / (2 * z)                    # fluxes.py:290
    energy, flux = [], []
    rows, row_u, row_E1, row_W, row_emit = [], [], [], [], []

    def single(el_energy, el_wgt):                                                   # fluxes.py:318-337
        with np.errstate(all="ignore"):
            ea_max = el_energy * (1 - np.power(ma / el_energy, 2))
            if ea_max <= ma:
                return False
            ea_rnd = np.power(10, usrc.uniform(np.log10(ma), np.log10(ea_max), n_samples))
            if boson_type == "pseudoscalar":
                mc_vol = np.log(10) * ea_rnd * (np.log10(ea_max) - np.log10(ma)) / n_samples
                diff_br = (ntarget_area_density * HBARC ** 2) * mc_vol \
                    * brem_dsigma_dea(ea_rnd, el_energy, g, ma, z)
            elif boson_type == "vector":
                x_rnd = ea_rnd / el_energy
                mc_vol = np.log(10) * x_rnd * (np.log10(ea_max / el_energy) - np.log10(ma / el_energy)) / n_samples
                diff_br = (ntarget_area_density * HBARC ** 2) * mc_vol * brem_dsigma_dx_vector(x_rnd, g, ma, z)
            else:
                raise UnboundLocalError("diff_br")   # the reference falls through to an unbound local
            energy.append(ea_rnd)
            flux.append(el_wgt * diff_br)
            return True

    if use_track_length:
        ep_min = max(ma, M_E)                                                        # fluxes.py:346
        for i, el in enumerate(electron_flux):
            if el[0] < ep_min:
                continue
            u = usrc.take()
            new_energy = ep_min + (el[0] - ep_min) * u                               # :350
            fl = (np.interp(el[0], electron_flux[:, 0], electron_flux[:, 1], left=0.0, right=0.0)
                  + np.interp(el[0], positron_flux[:, 0], positron_flux[:, 1], left=0.0, right=0.0))
            flux_weight = (el[0] - ep_min) * (fl * track_length_integrated_prob(el[0], new_energy, max_t))  # :351
            emitted = single(new_energy, flux_weight)
            rows.append(i); row_u.append(u); row_E1.append(new_energy); row_W.append(flux_weight)
            row_emit.append(emitted)
    else:
        for i, el in enumerate(electron_flux):                                       # fluxes.py:354-356
            emitted = single(el[0], el[1])
            rows.append(i); row_u.append(np.nan); row_E1.append(el[0]); row_W.append(el[1])
            row_emit.append(emitted)
    info = {"rows": np.array(rows, dtype=np.int64), "row_u": np.array(row_u), "row_E1": np.array(row_E1),
            "row_W": np.array(row_W), "row_emit": np.array(row_emit, dtype=bool)}
    if not energy:
        return np.zeros(0), np.zeros(0), info
    return np.concatenate(energy), np.concatenate(flux), info


# ---------------------------------------------------------------------------------------------------------
# fluxes.py:385-445  FluxResonanceIsotropic.simulate
# ---------------------------------------------------------------------------------------------------------
def resonance_peak(ge, ma, boson_type="pseudoscalar"):                               # fluxes.py:413-419
    if boson_type == "pseudoscalar":
        return 2 * pi * M_E * np.power(ge / ma, 2) / np.sqrt(1 - np.power(2 * M_E / ma, 2))
    elif boson_type == "vector":
        return 12 * pi * ge ** 2 * ALPHA / M_E / 6


def resonance_simulate(positron_flux, ma, g, target, n_samples, usrc, boson_type="pseudoscalar", max_t=5.0):
    """Draw order: `n_samples` for E+ (fluxes.py:437) then `n_samples` for t (:438).  One output sample."""
    positron_flux = np.asarray(positron_flux, dtype=np.float64)
    z = target.z[0]
    ntarget_area_density = target.rad_length * AVOGADRO / (target.z[0] + target.n[0])   # fluxes.py:397
    resonant_energy = -M_E + ma ** 2 / (2 * M_E)                                     # :427
    emax = max(positron_flux[:, 0])
    if resonant_energy + M_E < ma or resonant_energy < M_E or resonant_energy > emax:   # :428-435
        return np.zeros(0), np.zeros(0)
    e_rnd = usrc.uniform(resonant_energy, emax, n_samples)                           # :437
    t_rnd = usrc.uniform(0.0, max_t, n_samples)                                      # :438
    mc_vol = max_t * (emax - resonant_energy)                                        # :439
    att = np.interp(e_rnd, positron_flux[:, 0], positron_flux[:, 1], left=0.0, right=0.0) \
        * track_length_prob(e_rnd, resonant_energy, t_rnd)                           # :410-411
    attenuated_flux = mc_vol * np.sum(att) / n_samples                               # :441
    wgt = z * (ntarget_area_density * HBARC ** 2) * resonance_peak(g, ma, boson_type) * attenuated_flux   # :442
    return np.array([ma ** 2 / (2 * M_E)]), np.array([wgt])                          # :444-445


# ---------------------------------------------------------------------------------------------------------
# fluxes.py:554-600 FluxNuclearIsotropic, fluxes.py:926-991 FluxPi0Isotropic
# ---------------------------------------------------------------------------------------------------------
M_PI0 = 134.98                                                                       # constants.py:42


def nuclear_simulate(rates, ma, gann0, gann1, j=1, delta=0.0):
    energy, flux = [], []
    mu0, mu1 = 0.88, 4.71
    for r in np.asarray(rates, dtype=np.float64):
        if r[0] > ma:                                                                # fluxes.py:588
            br = ((j / (j + 1)) / (1 + delta ** 2) / pi / ALPHA) \
                * np.power(np.sqrt(r[0] ** 2 - ma ** 2) / r[0], 2 * j + 1) \
                * np.power((gann0 * r[2] + gann1) / ((mu0 - 0.5) * r[2] + (mu1 - r[3])), 2)      # :571-577
            energy.append(r[0]); flux.append(r[1] * br)
    return np.array(energy), np.array(flux)


def pi0_br(g, ma):                                                                   # fluxes.py:938-939
    return 2 * g ** 2 * (1 - np.power(ma / M_PI0, 2)) ** 3 / np.sqrt(4 * pi * ALPHA)


def pi0_simulate_flux(pi0_flux, ma, g, meson_rate, usrc, energy_cut=0.0, angle_cut=np.pi):
    pi0_flux = np.asarray(pi0_flux, dtype=np.float64)
    if ma > M_PI0:
        return np.zeros(0), np.zeros(0)
    p_cm = (M_PI0 ** 2 - ma ** 2) / (2 * M_PI0)
    e1_cm = np.sqrt(p_cm ** 2 + ma ** 2)
    cos_rnd = usrc.uniform(-1, 1, pi0_flux.shape[0])                                 # :953
    energy, flux = [], []
    for i, p in enumerate(pi0_flux):
        beta = p / np.sqrt(p ** 2 + M_PI0 ** 2)
        boost = np.power(1 - beta ** 2, -0.5)
        e_lab = boost * (e1_cm + beta * p_cm * cos_rnd[i])
        pz_lab = boost * (p_cm * cos_rnd[i] + beta * e1_cm)
        angle_lab = np.arccos(pz_lab / np.sqrt(e_lab ** 2 - ma ** 2))
        if e_lab < energy_cut or angle_lab > angle_cut:
            continue
        energy.append(e_lab); flux.append(meson_rate * pi0_br(g, ma) / pi0_flux.shape[0])
    return np.array(energy), np.array(flux)


def epem_to_alp_photon_dsigma_de(ea, ep, g=1.0, ma=1.0, z=1):                        # prod_xs.py:109-129
    s = 2 * M_E * (M_E + ep)
    ps_cm = np.sqrt((s - ma ** 2) ** 2 / (4 * s))
    pp_cm = np.sqrt(((s - 2 * M_E ** 2) ** 2 - 4 * M_E ** 4) / (4 * s))
    es_cm = np.sqrt(ps_cm ** 2 + ma ** 2)
    ep_cm = np.sqrt(pp_cm ** 2 + M_E ** 2)
    beta = np.sqrt(ep ** 2 - M_E ** 2) / (M_E + ep)
    gamma = np.power(1 - beta ** 2, -0.5)
    costheta = (ea / gamma - es_cm) / (beta * ps_cm)
    t = ma ** 2 + M_E ** 2 - 2 * (ep_cm * es_cm - pp_cm * ps_cm * costheta)
    m_st = z * 4 * pi * ALPHA * g ** 2 * (2 * s * M_E ** 4 + 2 * M_E ** 2 * (ma ** 4 - s * ma ** 2 - 2 * s * t)
                                          + s * (ma ** 4 - 2 * ma ** 2 * (s + t) + s ** 2 + 2 * s * t + 2 * t ** 2)) / s ** 2
    jacobian = 2 * pp_cm / gamma / beta
    return np.heaviside(ep - max((ma ** 2 - M_E ** 2) / (2 * M_E), M_E), 1.0) * jacobian * m_st / (16 * pi * (s - 4 * M_E ** 2) * s)


def pairgamma_simulate(positron_flux, ma, g, target, n_samples, usrc, target_radiation_length=6.76):
    """FluxPairAnnihilationGamma.simulate (fluxes.py:1090-1134)."""
    pf = np.asarray(positron_flux, dtype=np.float64)
    widths = pf[1:, 0] - pf[:-1, 0]
    centers = (pf[1:, 0] + pf[:-1, 0]) / 2
    z = target.z[0]
    ntad = target_radiation_length * AVOGADRO / (2 * z)
    ep_min = max((ma ** 2 - M_E ** 2) / (2 * M_E), M_E)
    sm_e, sm_w = [], []
    with np.errstate(all="ignore"):
        for i in range(centers.shape[0]):
            ep_lab = centers[i]
            if ep_lab < ep_min:
                continue
            t_depths = usrc.uniform(0.0, 5.0, n_samples)
            smeared = usrc.uniform(ep_min, ep_lab, n_samples)
            b = 4 / 3
            tlp = np.heaviside(ep_lab - smeared, 0.0) * abs(np.power(np.log(ep_lab / smeared), b * t_depths - 1)
                                                           / (ep_lab * _gamma(b * t_depths)))
            fw = np.interp(ep_lab, pf[:, 0], pf[:, 1], left=0.0, right=0.0) * tlp
            sm_w.extend(widths[i] * 5.0 * (ep_lab - ep_min) * fw / n_samples)
            sm_e.extend(smeared)
        energy, flux = [], []
        for ep, pos_flux in zip(sm_e, sm_w):
            beta = np.sqrt(ep ** 2 - M_E ** 2) / (M_E + ep)
            gamma = np.power(1 - beta ** 2, -0.5)
            s = 2 * M_E * (M_E + ep)
            ps_cm = np.sqrt((s - ma ** 2) ** 2 / (4 * s))
            es_cm = np.sqrt(ps_cm ** 2 + ma ** 2)
            es_min = gamma * (es_cm - ps_cm * beta)
            es_max = gamma * (es_cm + ps_cm * beta)
            ea = usrc.uniform(es_min, es_max, 1)
            mc_vol = 2 * gamma * beta * ps_cm
            cm = (ntad * HBARC ** 2) * epem_to_alp_photon_dsigma_de(ea, ep, g, ma, z)
            energy.extend(ea)
            flux.extend(cm * pos_flux * mc_vol)
    return np.array(energy), np.array(flux)


# ---------------------------------------------------------------------------------------------------------
# cross_section_mc.py:669-685  lorentz_boost for a boost along +-z (the only case the in-scope flux classes
# produce: parents and beams are on the z axis).  Row sums are evaluated in the order NumPy's 4x4 `mat @ pmu`
# produces for rows with two non-zero terms.
# ---------------------------------------------------------------------------------------------------------
def _boost_z(e, px, py, pz, vz):
    """Boost 4-vectors (arrays) by frame velocity (0,0,vz) -- restates cross_section_mc.py:676-685."""
    beta = np.sqrt(vz * vz)                                                          # Vector3.mag :60
    if beta == 0.0:
        return e, px, py, pz
    nz = vz / beta                                                                   # unit_vec :49-53
    gam = 1 / np.sqrt(1 - beta ** 2)                                                 # :680
    m03 = -gam * beta * nz
    m33 = 1 + (gam - 1) * nz * nz
    e_lab = gam * e + m03 * pz
    pz_lab = m03 * e + m33 * pz
    return e_lab, px + 0.0, py + 0.0, pz_lab


# ---------------------------------------------------------------------------------------------------------
# matrix_element.py:717-734  M2AssociatedProduction ; cross_section_mc.py:181-240  Scatter2to2MC
# ---------------------------------------------------------------------------------------------------------
def m2_associated(s, t, ma):
    u_prop = (M_E ** 2 + ma ** 2 - s - t)
    t_prop = (M_E ** 2 - t)
    Mt2 = -4 * (3 * M_E ** 4 - M_E ** 2 * (ma ** 2 + s) + t * (-ma ** 2 + s + t)) / t_prop ** 2
    Mu2 = -4 * (7 * M_E ** 4 + M_E ** 2 * (ma ** 2 - 3 * s - 4 * t) + t * (-ma ** 2 + s + t)) / u_prop ** 2
    MtMu = 8 * (-3 * M_E ** 4 + M_E ** 2 * (s - 2 * t) + t * (-ma ** 2 + s + t)) / (u_prop * t_prop)
    M2 = Mt2 + Mu2 + 2 * MtMu
    prefactor = (pi * ALPHA) * 1.0 ** 2
    p_pos_cm = np.sqrt((np.power(s - 2 * M_E ** 2, 2) - 4 * np.power(M_E, 4)) / (4 * s))
    e_pos_cm = np.sqrt(p_pos_cm ** 2 + M_E ** 2)
    tmax = M_E ** 2 - 2 * 0.01 * e_pos_cm * (e_pos_cm)
    return prefactor * M2 * np.heaviside(tmax - t, 0.0)


def pairann_row_valid(positron_flux, ma):                                            # fluxes.py:507
    ep = np.asarray(positron_flux, dtype=np.float64)[:, 0]
    return ~(ep < max((ma ** 2 - M_E ** 2) / (2 * M_E), M_E))


def pairann_simulate(positron_flux, ma, g, target, n_samples, usrc):
    """FluxPairAnnihilationIsotropic.simulate (fluxes.py:503-530).  Draw order per row above threshold:
    `n_samples` phi then `n_samples` theta (cross_section_mc.py:204-205) -- drawn BEFORE the s-threshold
    return of scatter_sim (:213-214)."""
    positron_flux = np.asarray(positron_flux, dtype=np.float64)
    z = target.z[0]
    ntarget_area_density = target.rad_length * AVOGADRO / (2 * z)                    # fluxes.py:484
    m1, m2, m3, m4 = M_E, M_E, ma, 0.0                                               # matrix_element.py:714
    energy, flux = [], []
    with np.errstate(all="ignore"):
        for el in positron_flux:
            ep_lab, pos_wgt = el[0], el[1]
            if ep_lab < max((ma ** 2 - M_E ** 2) / (2 * M_E), M_E):                  # fluxes.py:507
                continue
            p1z = np.sqrt(ep_lab ** 2 - M_E ** 2)                                    # :513
            phi = 2 * pi * usrc.take(n_samples)                                      # cross_section_mc.py:204
            theta = np.arccos(1 - 2 * usrc.take(n_samples))                          # :205
            e_in = ep_lab + M_E                                                      # :208-209
            vz = (p1z + 0.0) / e_in                                                  # :211
            # mass2 = dot(pmu**2, [1,-1,-1,-1])  (:110-111)
            s = float(np.dot(np.array([e_in, 0.0, 0.0, p1z]) ** 2, np.array([1, -1, -1, -1])))
            # cross_section_mc.py:213-214 (s < (m3+m4)^2 -> early return) is unreachable for rows that passed
            # the fluxes.py:507 threshold: s = 2 me^2 + 2 me E+ >= ma^2 + me^2 > ma^2.
            p1_cm = np.sqrt((np.power(s - m1 ** 2 - m2 ** 2, 2) - np.power(2 * m1 * m2, 2)) / (4 * s))
            p3_cm = np.sqrt((np.power(s - m3 ** 2 - m4 ** 2, 2) - np.power(2 * m3 * m4, 2)) / (4 * s))
            e1_cm = np.sqrt(p1_cm ** 2 + m1 ** 2)
            e3_cm = np.sqrt(p3_cm ** 2 + m3 ** 2)
            t_rnd = m1 ** 2 + m3 ** 2 + 2 * (p1_cm * p3_cm * np.cos(theta) - e1_cm * e3_cm)   # :221
            mc_volume = 4 * p1_cm * p3_cm / n_samples                                # :227
            dsdt = m2_associated(s, t_rnd, ma) / (16 * np.pi * (s - (m1 + m2) ** 2) * (s - (m1 - m2) ** 2))
            wts = mc_volume * dsdt                                                   # :228
            px = p3_cm * np.cos(phi) * np.sin(theta)                                 # :231-234
            py = p3_cm * np.sin(phi) * np.sin(theta)
            pz = p3_cm * np.cos(theta)
            # lorentz_boost(p3, -v_in) with v_in = (0,0,vz)                             :235
            ea_lab, _, _, _ = _boost_z(e3_cm * np.ones(n_samples), px, py, pz, -vz)
            energy.append(ea_lab)
            flux.append(pos_wgt * z * g ** 2 * ntarget_area_density * HBARC ** 2 * wts)   # fluxes.py:518
    if not energy:
        return np.zeros(0), np.zeros(0)
    return np.concatenate(energy), np.concatenate(flux)


# ---------------------------------------------------------------------------------------------------------
# Generic 2 -> 2 scattering MC: cross_section_mc.py:148-275 (Scatter2to2MC) with the matrix elements of
# matrix_element.py:629-734 and the atomic form factor form_factors.py:54-64
# ---------------------------------------------------------------------------------------------------------
def atomic_elastic_ff2(q, z):                                                        # form_factors.py:61-64
    t = q ** 2
    a = 184.15 * np.power(2.718, -1 / 2) * np.power(z, -1 / 3) / M_E
    return np.power(z * (t * a ** 2) / (1 + t * a ** 2), 2)


def _m2_compton_like(s, t, ma, z, coupling_product):                                 # matrix_element.py:638-645 / 659-666
    u = 2 * M_E ** 2 + ma ** 2 - s - t
    prefactor = z * 4 * pi * ALPHA * coupling_product ** 2
    Ms2 = prefactor * (M_E ** 4 + M_E ** 2 * (3 * ma ** 2 - 2 * s - t) + s * (s + t - ma ** 2)) / np.power(M_E ** 2 - s, 2)
    Mu2 = prefactor * (M_E ** 4 + M_E ** 2 * (3 * ma ** 2 - 2 * s - t) + s * (s + t - ma ** 2)) / np.power(M_E ** 2 - u, 2)
    MuMt = prefactor * (M_E ** 4 + M_E ** 2 * (3 * ma ** 2 - 2 * s - t) - (ma ** 2 - s) * (s + t)) / ((M_E ** 2 - s) * (M_E ** 2 - u))
    return Ms2 + Mu2 + 2 * MuMt


def m2_compton(s, t, ma, z, coupling_product=1.0):                                   # matrix_element.py:629-645
    return _m2_compton_like(s, t, ma, z, coupling_product)


def m2_inverse_compton(s, t, ma, z, coupling_product=1.0):                           # matrix_element.py:650-666
    return 2 * _m2_compton_like(s, t, ma, z, coupling_product)


def _m2_primakoff_like(s, t, ma, mN, z, prefactor):                                  # matrix_element.py:681-686 / 700-705
    propagator = np.power(t, 2)
    numerator = -t * (2 * mN ** 2 * (ma ** 2 - 2 * s - t) + 2 * mN ** 4 - 2 * ma ** 2 * (s + t) + ma ** 4 + 2 * s ** 2
                      + 2 * s * t + t ** 2)
    return atomic_elastic_ff2(np.sqrt(abs(t)), z) * prefactor * numerator / propagator


def m2_inverse_primakoff(s, t, ma, mN, z, coupling_product=1.0):                     # matrix_element.py:671-686
    return _m2_primakoff_like(s, t, ma, mN, z, 2 * coupling_product ** 2)


def m2_primakoff(s, t, ma, mN, z, coupling_product=1.0):                             # matrix_element.py:690-705
    return _m2_primakoff_like(s, t, ma, mN, z, coupling_product ** 2)


# (name, masses (m1, m2, m3, m4), callable(s, t)) of the matrix elements Scatter2to2MC is used with on this path
def matrix_element(model, ma, mN=0.0, z=1):
    if model == "inverse_primakoff":
        return (ma, mN, 0, mN), lambda s, t: m2_inverse_primakoff(s, t, ma, mN, z)
    if model == "primakoff":
        return (0, mN, ma, mN), lambda s, t: m2_primakoff(s, t, ma, mN, z)
    if model == "inverse_compton":
        return (ma, M_E, 0, M_E), lambda s, t: m2_inverse_compton(s, t, ma, z)
    if model == "compton":
        return (0, M_E, ma, M_E), lambda s, t: m2_compton(s, t, ma, z)
    if model == "associated_production":
        return (M_E, M_E, ma, 0.0), lambda s, t: m2_associated(s, t, ma)
    raise ValueError(model)


def boost_matrix(v):
    """The 4x4 matrix of lorentz_boost (cross_section_mc.py:676-685) for frame velocity v = (vx, vy, vz); None if
    beta == 0 (the reference returns the momentum unchanged)."""
    v = np.asarray(v, dtype=np.float64)
    beta = np.sqrt(np.dot(v, v))                                                     # Vector3.mag :59-60
    if beta == 0.0:
        return None
    n = v / beta if not beta < 1e-40 else np.zeros(3)                                # unit_vec :49-53
    g = 1 / np.sqrt(1 - beta ** 2)
    return np.array([[g, -g * beta * n[0], -g * beta * n[1], -g * beta * n[2]],
                     [-g * beta * n[0], 1 + (g - 1) * n[0] * n[0], (g - 1) * n[0] * n[1], (g - 1) * n[0] * n[2]],
                     [-g * beta * n[1], (g - 1) * n[1] * n[0], 1 + (g - 1) * n[1] * n[1], (g - 1) * n[1] * n[2]],
                     [-g * beta * n[2], (g - 1) * n[2] * n[0], (g - 1) * n[2] * n[1], 1 + (g - 1) * n[2] * n[2]]])


def scatter2to2_sim(masses, m2, p1, p2, n_samples, usrc, log_sampling=False):
    """Scatter2to2MC.scatter_sim (cross_section_mc.py:193-240) for incoming 4-vectors p1, p2 = (E, px, py, pz).
    Draw order: n phi, then n theta variates (or n logu with log_sampling) -- drawn before the s-threshold return.

    Returns None below threshold (the reference leaves its lists empty), else a dict with
      t[n], weights[n] (dsigma_dcos_cm_wgts), p3_cm[n,4], p3_lab[n,4], p4_lab[n,4],
      cosines[n] (get_cosine_lab_weights), e3_lab[n], e4_lab[n], and the 3-velocity arrays."""
    m1, m2_, m3, m4 = masses
    p1 = np.asarray(p1, dtype=np.float64)
    p2 = np.asarray(p2, dtype=np.float64)
    n = int(n_samples)
    with np.errstate(all="ignore"):
        if log_sampling:
            phi = 2 * pi * usrc.take(n)                                              # :199
            logu = usrc.uniform(-12, 0, n)                                           # :200
            theta = np.arccos(1 - 2 * np.exp(logu))                                  # :201
        else:
            phi = 2 * pi * usrc.take(n)                                              # :203
            theta = np.arccos(1 - 2 * usrc.take(n))                                  # :204
        cm = p1 + p2                                                                 # :207 (component-wise LorentzVector add)
        e_in = cm[0]
        v_in = np.array([cm[1] / e_in, cm[2] / e_in, cm[3] / e_in])                  # :210
        s = float(np.dot(cm ** 2, np.array([1, -1, -1, -1])))                        # mass2 :110-111
        if s < (m3 + m4) ** 2:                                                       # :212-213
            return None
        p1_cm = np.sqrt((np.power(s - m1 ** 2 - m2_ ** 2, 2) - np.power(2 * m1 * m2_, 2)) / (4 * s))   # :187-188
        p3_cm = np.sqrt((np.power(s - m3 ** 2 - m4 ** 2, 2) - np.power(2 * m3 * m4, 2)) / (4 * s))     # :190-191
        e1_cm = np.sqrt(p1_cm ** 2 + m1 ** 2)
        e3_cm = np.sqrt(p3_cm ** 2 + m3 ** 2)
        t_rnd = m1 ** 2 + m3 ** 2 + 2 * (p1_cm * p3_cm * np.cos(theta) - e1_cm * e3_cm)                # :220
        if log_sampling:
            mc_volume = np.exp(logu) * 2 * p1_cm * p3_cm / n                         # :224
        else:
            mc_volume = 4 * p1_cm * p3_cm / n                                        # :226
        dsdt = m2(s, t_rnd) / (16 * np.pi * (s - (m1 + m2_) ** 2) * (s - (m1 - m2_) ** 2))             # :184-185
        wts = mc_volume * dsdt                                                       # :227
        p3cm = np.stack([e3_cm * np.ones(n), p3_cm * np.cos(phi) * np.sin(theta), p3_cm * np.sin(phi) * np.sin(theta),
                         p3_cm * np.cos(theta)], axis=1)                             # :230-233
        mat = boost_matrix(-v_in)                                                    # :234  lorentz_boost(p3, -v_in)
        if mat is None:
            p3lab = p3cm.copy()
        else:
            p3lab = np.stack([((mat[r, 0] * p3cm[:, 0] + mat[r, 1] * p3cm[:, 1]) + mat[r, 2] * p3cm[:, 2]) + mat[r, 3] * p3cm[:, 3]
                              for r in range(4)], axis=1)
        p4lab = (p1 + p2)[None, :] - p3lab                                           # :238
        mom3 = np.sqrt((p3lab[:, 1] * p3lab[:, 1] + p3lab[:, 2] * p3lab[:, 2]) + p3lab[:, 3] * p3lab[:, 3])   # Vector3.mag
        return {"s": s, "t": t_rnd, "weights": wts, "p3_cm": p3cm, "p3_lab": p3lab, "p4_lab": p4lab,
                "cosines": p3lab[:, 3] / mom3,                                       # :257 LorentzVector.cosine :119-120
                "e3_lab": p3lab[:, 0], "e4_lab": p4lab[:, 0],                        # :268, :272
                "v3_cm": p3cm[:, 1:] / p3cm[:, :1], "v3_lab": p3lab[:, 1:] / p3lab[:, :1],   # get_3velocity :142-143
                "v4_lab": p4lab[:, 1:] / p4lab[:, :1]}


# ---------------------------------------------------------------------------------------------------------
# generators.py:41-46, 80-92  event weights
# ---------------------------------------------------------------------------------------------------------
def decays(axion_energy, decay_axion_weight, days_exposure, threshold):              # generators.py:88-92
    """`days*S_PER_DAY*w` keeps w's dtype (float32 attribute -> float32 product under NEP 50 for a Python
    `days`), the heaviside factor then promotes to float64."""
    e = np.asarray(axion_energy, dtype=np.float64)
    w = days_exposure * S_PER_DAY * decay_axion_weight * np.heaviside(e - threshold, 1.0)
    return w, np.sum(w)


def compton_events(axion_energy, scatter_axion_weight, ge, ma, ntargets, det_area, det_z,
                   days_exposure, threshold):                                        # generators.py:41-46
    e = np.asarray(axion_energy, dtype=np.float64)
    w = days_exposure * S_PER_DAY * (ntargets / det_area) \
        * icompton_sigma(e, ma, ge, det_z) \
        * METER_BY_MEV ** 2 * scatter_axion_weight * np.heaviside(e - threshold, 1.0)
    return w, np.sum(w)


def inverse_primakoff_events(axion_energy, scatter_axion_weight, gagamma, ma, ntargets, det_area, det_z,
                             days_exposure, threshold=0.0):                          # generators.py:80-86
    e = np.asarray(axion_energy, dtype=np.float64)
    w = days_exposure * S_PER_DAY * (ntargets / det_area) \
        * iprimakoff_sigma(e, gagamma, ma, det_z) \
        * METER_BY_MEV ** 2 * scatter_axion_weight * np.heaviside(e - threshold, 1.0)
    return w, np.sum(w)


class ALPPairProdutionCrossSection:                                                  # photon_xs.py:95-125
    def __init__(self, material, ma=0.1):
        ext = ["100keV", "400keV", "500keV", "600keV", "700keV", "800keV", "900keV", "1MeV", "1.1MeV"]
        cp = [0.1, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1]
        idx = np.clip(int(np.interp(ma, cp, np.arange(0, 9, 1))), a_min=0, a_max=9)
        self.argon_rescale = np.power(material.z[0] / 18, 2)
        self.xs_dim = 1e-24
        self.xs_data = np.genfromtxt(os.path.join(_DATA_DIR, "alp_pair_production", "pair_production_xs_table_" + ext[idx] + ".txt"))

    def sigma_mev(self, E):
        with np.errstate(all="ignore"):
            return np.heaviside(E - 2 * M_E, 0.0) * np.power(10, np.interp(
                np.log10(E), np.log10(self.xs_data[:, 0]),
                np.log10(self.argon_rescale * self.xs_dim * self.xs_data[:, 1] / MEV2_CM2), left=-np.inf))


def pair_production_events(axion_energy, scatter_axion_weight, pair_xs, ge, ntargets, det_area, days_exposure,
                           threshold):                                               # generators.py:33-39
    e = np.asarray(axion_energy, dtype=np.float64)
    w = days_exposure * S_PER_DAY * (ntargets / det_area) \
        * np.power(ge, 2) * pair_xs.sigma_mev(e) \
        * METER_BY_MEV ** 2 * scatter_axion_weight * np.heaviside(e - threshold, 1.0)
    return w, np.sum(w)


# ---------------------------------------------------------------------------------------------------------
# generators.py:94-128 + cross_section_mc.py:510-544  simulate_decay_4vectors (forward parents, theta = 0)
# ---------------------------------------------------------------------------------------------------------
def decay_4vectors(axion_energy, decay_axion_weight, ma, days_exposure, n_samples, usrc):
    """Draw order: `n_flux` parent phis (generators.py:100, drawn but irrelevant at theta=0), then per parent
    `n_samples` phi1 and `n_samples` theta1 (cross_section_mc.py:529-530).

    Returns p1[N,4], p2[N,4] (E,px,py,pz), w[N] float32, N = n_flux*n_samples, parent-major order."""
    e_all = np.asarray(axion_energy, dtype=np.float64)
    n_flux = e_all.shape[0]
    event_weight = days_exposure * S_PER_DAY * decay_axion_weight / n_samples        # generators.py:97
    phis = usrc.uniform(0.0, 2 * pi, n_flux)                                         # :100 (unused at theta=0)
    p1 = np.empty((n_flux * n_samples, 4))
    p2 = np.empty((n_flux * n_samples, 4))
    w = np.empty(n_flux * n_samples, dtype=np.asarray(event_weight).dtype)
    m1 = m2 = 0.0
    with np.errstate(all="ignore"):
        for i in range(n_flux):
            ea = e_all[i]
            pa = np.sqrt(ea ** 2 - ma ** 2)                                          # :118
            # theta=0: px = pa*cos(phi)*sin(0) = +-0, py = +-0, pz = pa*cos(0) = pa  (:120)
            pz_parent = pa * np.cos(0.0)
            mp = np.sqrt(ea ** 2 - pz_parent ** 2)                                   # LorentzVector.mass() :113-114
            p_cm = np.power((mp ** 2 - (m2 - m1) ** 2) * (mp ** 2 - (m2 + m1) ** 2), 0.5) / (2 * mp)   # :524
            e1_cm = np.sqrt(p_cm ** 2 + m1 ** 2)
            e2_cm = np.sqrt(p_cm ** 2 + m2 ** 2)
            phi1 = 2 * pi * usrc.take(n_samples)                                     # :529
            th1 = np.arccos(1 - 2 * usrc.take(n_samples))                            # :530
            vz = -(pz_parent / ea)                                                   # :532  v_in = -velocity
            px = p_cm * np.cos(phi1) * np.sin(th1)
            py = p_cm * np.sin(phi1) * np.sin(th1)
            pz = p_cm * np.cos(th1)
            sl = slice(i * n_samples, (i + 1) * n_samples)
            a = _boost_z(e1_cm * np.ones(n_samples), px, py, pz, vz)
            b = _boost_z(e2_cm * np.ones(n_samples), -px, -py, -pz, vz)
            p1[sl] = np.stack(a, axis=1)
            p2[sl] = np.stack(b, axis=1)
            w[sl] = event_weight[i] * np.ones(n_samples)                             # generators.py:126
    return p1, p2, w


# ---------------------------------------------------------------------------------------------------------
# General Decay2Body (cross_section_mc.py:496-544) with the dense lorentz_boost (:669-685); row sums evaluated
# left to right (NumPy/BLAS may associate differently at the 1-ulp level)
# ---------------------------------------------------------------------------------------------------------
def decay2body_general(parent_p4, m1, m2, n_samples, usrc):
    """parent_p4 = (E, px, py, pz).  Returns p1[n,4], p2[n,4] lab 4-vectors.  Draws: n phi then n theta."""
    e, px, py, pz = (float(x) for x in parent_p4)
    with np.errstate(all="ignore"):
        mp = np.sqrt(((e * e - px * px) - py * py) - pz * pz)
        p_cm = np.power((mp ** 2 - (m2 - m1) ** 2) * (mp ** 2 - (m2 + m1) ** 2), 0.5) / (2 * mp)
        e1 = np.sqrt(p_cm ** 2 + m1 ** 2)
        e2 = np.sqrt(p_cm ** 2 + m2 ** 2)
        phi = 2 * pi * usrc.take(n_samples)
        th = np.arccos(1 - 2 * usrc.take(n_samples))
        v = np.array([-(px / e), -(py / e), -(pz / e)])
        beta = np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
        cx, cy, cz = p_cm * np.cos(phi) * np.sin(th), p_cm * np.sin(phi) * np.sin(th), p_cm * np.cos(th)
        a = np.stack([e1 * np.ones(n_samples), cx, cy, cz], axis=1)
        b = np.stack([e2 * np.ones(n_samples), -cx, -cy, -cz], axis=1)
        if beta == 0.0:
            return a, b
        n = v / beta if not beta < 1e-40 else np.zeros(3)
        g = 1 / np.sqrt(1 - beta ** 2)
        mat = np.array([[g, -g * beta * n[0], -g * beta * n[1], -g * beta * n[2]],
                        [-g * beta * n[0], 1 + (g - 1) * n[0] * n[0], (g - 1) * n[0] * n[1], (g - 1) * n[0] * n[2]],
                        [-g * beta * n[1], (g - 1) * n[1] * n[0], 1 + (g - 1) * n[1] * n[1], (g - 1) * n[1] * n[2]],
                        [-g * beta * n[2], (g - 1) * n[2] * n[0], (g - 1) * n[2] * n[1], 1 + (g - 1) * n[2] * n[2]]])

        def apply(p):
            return np.stack([((mat[r, 0] * p[:, 0] + mat[r, 1] * p[:, 1]) + mat[r, 2] * p[:, 2]) + mat[r, 3] * p[:, 3]
                             for r in range(4)], axis=1)
        return apply(a), apply(b)


def neutral_meson_simulate(meson_flux, flux_weight, ma, g, m_meson, n_samples, det_dist, det_area, usrc,
                           apply_angle_cut=False, off_axis_angle=0.0):
    """FluxNeutralMeson2BodyDecay.simulate (fluxes.py:1024-1046).  Returns (E, flux, theta)."""
    det_sa = np.arctan(np.sqrt(det_area / pi) / det_dist)
    phi_range = 2 * det_sa if off_axis_angle > 0.0 else 2 * pi
    br = 0.0 if ma > m_meson else 2 * g ** 2 * (1 - np.power(ma / m_meson, 2)) ** 3 / (4 * pi * ALPHA)
    E, F, T = [], [], []
    for m in np.asarray(meson_flux, dtype=np.float64):
        p1, _ = decay2body_general((m[3], m[0], m[1], m[2]), ma, 0.0, n_samples, usrc)
        en = p1[:, 0]
        th = np.arccos(p1[:, 3] / np.sqrt((p1[:, 1] ** 2 + p1[:, 2] ** 2) + p1[:, 3] ** 2))
        w = br * (np.ones(n_samples) / n_samples) * flux_weight
        if apply_angle_cut:
            mask = (th < det_sa + off_axis_angle) * (th > off_axis_angle - det_sa)
            en, w, th = en[mask], w[mask] * phi_range / (2 * pi), th[mask]
        else:
            w = det_area * w / (4 * pi * det_dist ** 2)
        E.append(en); F.append(w); T.append(th)
    return np.concatenate(E), np.concatenate(F), np.concatenate(T)


# ---------------------------------------------------------------------------------------------------------
# np.histogram(E, bins=edges, weights=w) semantics used for the binned spectra
# ---------------------------------------------------------------------------------------------------------
def histogram(x, w, edges):
    """np.histogram(x, bins=edges, weights=w)[0] semantics -- bin k = [edges[k], edges[k+1]), last bin closed, values
    outside the edges and NaN dropped -- evaluated as a direct per-bin sum (np.bincount).  np.histogram itself forms
    the bins as differences of a cumulative sum over the sorted weights, which cancels catastrophically when weights
    span many decades (decay weights do: 1e-6 .. 1e-22); tests/test_oracle_golden.py checks both agree on
    well-conditioned input."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    edges = np.asarray(edges, dtype=np.float64)
    nb = edges.shape[0] - 1
    idx = np.searchsorted(edges, x, side="right") - 1
    idx[x == edges[-1]] = nb - 1
    ok = (x >= edges[0]) & (x <= edges[-1])
    return np.bincount(idx[ok], weights=w[ok], minlength=nb)[:nb]


# ---------------------------------------------------------------------------------------------------------
# Charged-meson 3-body decay fluxes  M -> l nu X  (fluxes.py:605-922, matrix_element.py:50-68, 419-518)
# ---------------------------------------------------------------------------------------------------------
M_MU = 105.658369                     # constants.py:34
M_PI = 139.57                         # constants.py:41
M_K = 493.677                         # constants.py:53
G_F = 1.16638e-11                     # constants.py:73
CABIBBO = 0.9743                      # constants.py:74
V_UD = 0.97373                        # constants.py:78
V_US = 0.2243                         # constants.py:79
F_PI = 130.2                          # constants.py:91
F_K = 155.7                           # constants.py:92
PION_LIFETIME = 2.6e-8                # constants.py:108
PION_WIDTH = 2.524e-14                # constants.py:112
KAON_WIDTH = 5.317e-14                # constants.py:114
MESON_PARAMS = {"pion": (M_PI, V_UD, F_PI, PION_WIDTH), "kaon": (M_K, V_US, F_K, KAON_WIDTH)}   # fluxes.py:612-615


def W_ff(g, mf, ma):                                                                 # decay.py:18-21
    return g ** 2 * ma * np.sqrt(1 - (2 * mf / ma) ** 2) / (8 * pi) if 1 - 4 * (mf / ma) ** 2 > 0 else 0.0


def heaviside(x, h0):
    return np.heaviside(x, h0)


def m3_scalar_products(M, m1, m2, m3, m122, m232):                                   # matrix_element.py:60-68
    sp03 = (M ** 2 + m3 ** 2 - m122) / 2
    sp01 = (M ** 2 + m1 ** 2 - m232) / 2
    sp02 = (m122 + m232 - m1 ** 2 - m3 ** 2) / 2
    sp13 = (M ** 2 + m2 ** 2 - m122 - m232) / 2
    sp23 = (m232 - m2 ** 2 - m3 ** 2) / 2
    sp12 = (m122 - m1 ** 2 - m2 ** 2) / 2
    return sp01, sp02, sp03, sp12, sp13, sp23


def m2_meson3body(m122, m232, model, M, m1, m3, fM, ckm, c0=0.0, coupling=1.0, abd=(0, 0, 0)):
    """M2Meson3BodyDecay.__call__ (matrix_element.py:443-518); m2 (neutrino) = 0."""
    power, sqrt = np.power, np.sqrt
    lp, pq, kp, lq, kl, kq = m3_scalar_products(M, m1, 0.0, m3, m122, m232)
    if model in ("scalar_ib1", "pseudoscalar_ib1"):                                  # :445-464
        e3star = (M ** 2 - m122 - m3 ** 2) / (2 * sqrt(m122))
        ev = (m122 + m232 - m1 ** 2 - m3 ** 2) / (2 * M)
        emu = (M ** 2 - m232 + m1 ** 2) / (2 * M)
        q2 = M ** 2 - 2 * M * ev
        prefactor = heaviside(e3star - m3, 0.0) * (coupling * G_F * fM * ckm / (q2 - m1 ** 2)) ** 2
        head = (2 * M * emu * q2 * (q2 - m1 ** 2) - (q2 ** 2 - (m1 * M) ** 2) * (q2 + m1 ** 2 - m3 ** 2))
        tail = (2 * q2 * m1 ** 2 * (M ** 2 - q2))
        return prefactor * (head + tail) if model == "scalar_ib1" else prefactor * (head - tail)
    if model == "vector_ib1":                                                        # :465-472
        q2 = M ** 2 - 2 * M * (m122 + m232 - m1 ** 2 - m3 ** 2) / (2 * M)
        prefactor = 8 * power(G_F * fM * ckm / (q2 - m1 ** 2) / m3, 2)
        kl, kq, kp, lq, lp, pq = m3_scalar_products(M, m1, 0.0, m3, m122, m232)
        cr = cl = coupling
        return -prefactor * ((power(cr * M * m1, 2) - power(cl * q2, 2)) * (lq * m3 ** 2 + 2 * lp * pq)
                             - 2 * cr * m1 ** 2 * kq * (cr * m3 ** 2 * kl + 2 * cr * kp * lp - 3 * cl * q2 * m3 ** 2))
    if model == "scalar_ib2":                                                        # :473-474
        return 0.0 * m232
    if model == "vector_ib2":                                                        # :475-478
        prefactor = 2 * power(coupling * G_F * fM * m1, 2)
        return -prefactor * ((m1 ** 2 - m122) * ((M ** 2 - m122 + m3 ** 2) ** 2 - 4 * M ** 2 * m3 ** 2)) \
            / (m3 ** 2 * (M ** 2 - m122) ** 2) + 0.0 * m232
    if model == "vector_ib9":                                                        # :479-482, 500-518
        a, b, d = abd
        mV, ml, F, g = m3, m1, 1, coupling
        res = 1/2 * mV**-2 * (b**2 * (M**2 + -1 * m122 + -1 * m232)**3 * (m232 + -1 * mV**2) + 2 * b * d * ml**2 * (m232 + -1 * mV**2)**3 + 4 * b * (a
            * m122 + (a + 2 * d * m122) * ml**2) * mV**2 * (-1 * m232 + mV**2) + -4 * (m122 + -1 * ml**2) * mV**2 * (-1 * a**2 + 2 * a * d * ml**2
            + d**2 * m122 * ml**2 + (-1 * b**2 * m122 + F**2 * (m122 + ml**2)) * mV**2)
            + (m232 + -1 * mV**2)**2 * (4 * a * d * ml**2 + d**2 * ml**2 * (m122 + -1 * ml**2) + (b**2 * (-1 * m122 + ml**2)
            + 2 * F**2 * (m122 + ml**2)) * mV**2) + 2 * (-1 * M**2 + m122 + m232)**2 * (b * (2 * a + d * ml**2)
            * (m232 + -1 * mV**2) + b**2 * (m232 + -1 * mV**2)**2 + 1 / 2 * (m122 + -1 * ml**2) * (d**2 * ml**2 + -1 *
            (b**2 + -2 * F**2) * mV**2)) + (M**2 + -1 * m122 + -1 * m232) * (4 * a * b * (-1 * m122 + ml**2) * mV**2
            + 4 * b * (a + d * ml**2) * (m232 + -1 * mV**2)**2 + b**2 * (m232 + -1 * mV**2)**3 + 2 * (m232 + -1
            * mV**2) * (2 * a**2 + 2 * a * d * ml**2 + d**2 * ml**2 * (m122 + -1 * ml**2) + b**2 * (-3 * m122 + ml**2) * mV**2)))
        return g**2 * res * G_F**2 * fM**2
    if model == "sd":                                                                # :483-490
        prefactor = 8 * power(coupling * G_F * CABIBBO / sqrt(2) / M, 2)
        fv = 0.0265
        fa = 0.58 * fv
        return prefactor * (2 * kl * (power(fa - fv, 2) * kp * pq - M ** 2 * (fa ** 2 + fv ** 2) * kq)
                            + m3 ** 2 * (power(fa * M, 2) * lq - 2 * fv ** 2 * lp * pq)
                            + 2 * power(fa + fv, 2) * kp * kq * lp)
    if model == "vector_contact":                                                    # :491-496
        denom = power((c0 + 1) * m3 * (m3 ** 2 - kp), 2)
        numerator1 = 8 * m3 ** 2 * (2 * c0 * (kq * lp) * (kp - m3 ** 2)
                                    - lq * ((1 - c0 ** 2) * kp ** 2 - 2 * m3 ** 2 * kp + power(c0 * M * m3, 2) + m3 ** 4))
        numerator2 = 16 * kl * (kq * ((c0 + 1) * kp + m3 * (c0 * M - m3)) * (m3 * (c0 * M + m3) - (c0 + 1) * kp)
                                + c0 * m3 ** 2 * pq * (kp - m3 ** 2))
        return -power(coupling * G_F * fM * ckm / sqrt(2), 2) * (numerator1 + numerator2) / denom
    return 0.0 * m232                                                                # MatrixElementDecay3.__call__ :57-58


def meson3_dgamma_dea(Ea, ml, ma, mm, fM, ckm, model, c0, coupling, abd=(0, 0, 0)):
    """dGammadEa (fluxes.py:639-661 / 841-863): Dalitz integral over m_23^2 with scipy quad, like the reference.
    Only the two hard-wired lepton masses give a non-zero result (:649-661)."""
    sqrt, power = np.sqrt, np.power
    m212 = mm ** 2 + ma ** 2 - 2 * mm * Ea
    e2star = (m212 - ml ** 2) / (2 * sqrt(m212))
    e3star = (mm ** 2 - m212 - ma ** 2) / (2 * sqrt(m212))
    if ma > e3star:
        return 0.0
    m223Max = (e2star + e3star) ** 2 - (sqrt(e2star ** 2) - sqrt(e3star ** 2 - ma ** 2)) ** 2
    m223Min = (e2star + e3star) ** 2 - (sqrt(e2star ** 2) + sqrt(e3star ** 2 - ma ** 2)) ** 2
    if ml != M_MU and ml != M_E:
        return 0.0

    def matrix_element2(m223):
        return m2_meson3body(m212, m223, model, mm, ml, ma, fM, ckm, c0=c0, coupling=coupling, abd=abd)

    return (2 * mm) / (32 * power(2 * pi * mm, 3)) * quad(matrix_element2, m223Min, m223Max)[0]


def charged_meson_isotropic_simulate(meson_flux, ma, coupling, meson_type, model, n_samples, c0, lepton_masses, abd, usrc):
    """FluxChargedMeson3BodyIsotropic.simulate (fluxes.py:877-915).  Draw order per (meson row, allowed lepton):
    n energies, then n cosines."""
    mm, ckm, fM, total_width = MESON_PARAMS[meson_type]
    energy, flux = [], []
    for p in meson_flux:
        for ml in lepton_masses:
            if ma > mm - ml:                                                         # :908
                continue
            ea_min = ma
            ea_max = (mm ** 2 + ma ** 2 - ml ** 2) / (2 * mm)
            energies = usrc.uniform(ea_min, ea_max, n_samples)                       # :882
            momenta = np.sqrt(energies ** 2 - ma ** 2)
            cosines = usrc.uniform(-1, 1, n_samples)                                 # :884
            pz = momenta * cosines
            beta = p[0] / np.sqrt(p[0] ** 2 + mm ** 2)                               # :888
            boost = np.power(1 - beta ** 2, -0.5)
            e_lab = boost * (energies + beta * pz)                                   # :890
            mc_vol = ea_max - ea_min
            weights = np.array([p[1] * mc_vol * meson3_dgamma_dea(ea, ml, ma, mm, fM, ckm, model, c0, coupling, abd)
                                / total_width / n_samples for ea in energies])      # :898-899
            energy.extend(e_lab); flux.extend(weights)
    return np.array(energy), np.array(flux)


def charged_meson_decay_simulate(meson_flux, ma, coupling, meson_type, model, n_samples, c0, lepton_masses,
                                 energy_cut, det_dist, det_length, usrc, cut_on_solid_angle=True):
    """FluxChargedMeson3BodyDecay.simulate / simulate_single, single-process branch (fluxes.py:669-741).
    Draw order per meson row: one exponential decay position (scipy expon.rvs -> legacy standard_exponential =
    -log(1 - u)); then per allowed lepton (if decay_pos <= 52): n energies, n cosines, n azimuths.
    Returns dict of arrays: energy, cosine, flux, angle, solid_angle, decay_pos."""
    mm, ckm, fM, total_width = MESON_PARAMS[meson_type]
    cos, sin, arccos, arctan, sqrt, power = np.cos, np.sin, np.arccos, np.arctan, np.sqrt, np.power
    out = {k: [] for k in ("energy", "cosine", "flux", "angle", "solid_angle", "decay_pos")}
    positions = []
    for p in meson_flux:
        scale = p[0] * PION_LIFETIME * 0.01 * C_LIGHT / M_PI                         # :740
        x = -np.log(1.0 - usrc.take()) * scale + 0.0
        positions.append(x)
        for ml in lepton_masses:
            if ma > mm - ml:                                                         # :747
                continue
            if x > 52.0:                                                             # :671
                continue
            ea_min = ma
            ea_max = (mm ** 2 + ma ** 2 - ml ** 2) / (2 * mm)
            beta = p[0] / sqrt(p[0] ** 2 + mm ** 2)                                  # :682
            boost = power(1 - beta ** 2, -0.5)
            solid_angle_cosine = cos(arctan(det_length / (det_dist - x)))            # :686
            min_cm_cos = cos(min(boost * arccos(solid_angle_cosine), pi))            # :687
            energies = usrc.uniform(ea_min, ea_max, n_samples)                       # :689
            momenta = sqrt(energies ** 2 - ma ** 2)
            cosines = usrc.uniform(min_cm_cos, 1, n_samples)                         # :691
            pz = momenta * cosines
            e_lab = boost * (energies + beta * pz)                                   # :695
            pz_lab = boost * (pz + beta * energies)
            theta_lab = arccos(pz_lab / sqrt(e_lab ** 2 - ma ** 2))
            phi_lab = usrc.uniform(0, 2 * pi, n_samples)                             # :698
            thetas_z = arccos(cos(theta_lab) * cos(p[1]) + cos(phi_lab) * sin(theta_lab) * sin(p[1]))   # :701
            mc_vol_lab = (1.0 - min_cm_cos) * boost * beta * sqrt(e_lab ** 2 - ma ** 2)               # :705
            weights = np.array([p[2] * mc_vol_lab[i] * meson3_dgamma_dea(energies[i], ml, ma, mm, fM, ckm, model, c0, coupling)
                                / total_width / n_samples for i in range(energies.shape[0])])         # :708-709
            for i in range(n_samples):                                               # :715-725
                accept = heaviside(cos(theta_lab[i]) - solid_angle_cosine, 0.0)
                if accept == 0.0 and cut_on_solid_angle:
                    continue
                out["energy"].append(e_lab[i])
                out["cosine"].append(cos(theta_lab[i]))
                out["flux"].append(weights[i] * heaviside(e_lab[i] - energy_cut, 1.0))
                out["angle"].append(thetas_z[i])
                out["solid_angle"].append(solid_angle_cosine)
                out["decay_pos"].append(x)
    res = {k: np.array(v, dtype=np.float64) for k, v in out.items()}
    res["positions"] = np.array(positions)
    return res


# ---------------------------------------------------------------------------------------------------------
# FluxAxionMesonMixing (fluxes.py:1148-1240)
# ---------------------------------------------------------------------------------------------------------
M_ETA = 547.862                       # constants.py:43
M_ETA_PRIME = 957.78                  # constants.py:44


def meson_mixing_photon_coupling(ma, f_a, c1=1.0, c2=1.0, c3=1.0):                  # fluxes.py:1176-1185
    c_gamma = c2 + (5/3)*c1 + c3 * (
        -1.92
        + (1/3) * (ma**2 / (ma**2 - M_PI0**2))
        + (8/9) * ((ma**2 - (4/9)*M_PI0**2) / (ma**2 - M_ETA**2))
        + (7/9) * ((ma**2 - (16/9)*M_PI0**2) / (ma**2 - M_ETA_PRIME**2))
    )
    return c_gamma * ALPHA / (2 * pi * f_a)


def W_agg_hadronic(f_a, ma):                                                         # decay.py:67-72
    data = np.genfromtxt(os.path.join(_DATA_DIR, "gluon_dominance_widths_nogamma.txt"))
    return np.interp(ma*1e-3, data[:, 0], data[:, 1]) * 1e-9 / (32 * pi**2 * f_a / 1e6)**2


def meson_mixing_simulate(meson_flux, ma, f_a, meson_mass, meson_type, total_pot, mesons_per_pot, det_dist, det_area):
    """FluxAxionMesonMixing.simulate (fluxes.py:1202-1224): (energy, angle, flux)."""
    if meson_type == "Pi0":
        mixing = (1/6) * (F_PI/f_a) * (ma**2 / (ma**2 - M_PI0**2))
    else:
        mixing = (1/np.sqrt(6)) * (F_PI/f_a) * ((ma**2 - 4*M_PI0**2 / 9) / (ma**2 - M_ETA**2))
    psf = 1.0 if ma <= meson_mass else np.power(ma / meson_mass, -1.6)
    det_sa = np.arctan(np.sqrt(det_area / pi) / det_dist)                            # fluxes.py:40-41
    e_, th_, f_ = [], [], []
    for m in meson_flux:
        if m[0] < ma:
            continue
        angle = np.arccos(m[3]/np.sqrt(m[1]**2 + m[2]**2 + m[3]**2))
        if angle > det_sa:
            continue
        e_.append(m[0]); th_.append(angle)
        f_.append(total_pot * mesons_per_pot * (1.0/meson_flux.shape[0]) * psf * mixing**2)
    return np.array(e_), np.array(th_), np.array(f_)

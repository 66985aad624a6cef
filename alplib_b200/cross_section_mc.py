"""
Generic 2 -> 2 scattering Monte Carlo -- drop-in for the reference's `Scatter2to2MC` (cross_section_mc.py:148-275,
usage README.md:144-158), plus the value types `Vector3`, `LorentzVector` and `lorentz_boost` it is used with.

`scatter_sim()` draws the CM angles, evaluates t, the matrix element and the Monte-Carlo weights, builds the outgoing
4-vectors in the CM frame, boosts them to the frame of (p1, p2) and forms p4 = p1 + p2 - p3 -- all in ONE CUDA kernel
(libalpk `alpk_scatter2to2`, one thread per sample).  The per-call scalars (s, CM momenta and energies, the boost matrix)
are evaluated on the host with the reference's own expressions.  Results stay on the GPU; the reference's list attributes
(`p3_lab_4vectors`, ...) are lazily materialised `LorentzVectorArray` / arrays.

Random variates: Philox4x32-10 keyed by (seed, stream tag 8, call counter, sample); `seed=None` draws the key from the global
NumPy RNG (so np.random.seed pins it).  `scatter_sim(uniforms=...)` injects U[0,1) variates in the reference's draw
order (n phi, then n theta / logu) for parity tests.
"""
import ctypes as C

import numpy as np
import torch
from numpy import pi

from . import _native as N
from .matrix_element import MatrixElement2
from .runtime import Runtime
from .vectors import LorentzVector, LorentzVectorArray, Vector3, planes_to_host_rows

__all__ = ["Scatter2to2MC", "LorentzVector", "Vector3", "lorentz_boost"]


def _boost_matrix(v: Vector3):
    """The matrix of `lorentz_boost` (cross_section_mc.py:676-684); None when beta == 0."""
    n = v.unit_vec().vec
    beta = v.mag()
    if beta == 0.0:
        return None
    gamma = 1/np.sqrt(1-beta**2)
    return np.array([[gamma, -gamma*beta*n[0], -gamma*beta*n[1], -gamma*beta*n[2]],
                     [-gamma*beta*n[0], 1+(gamma-1)*n[0]*n[0], (gamma-1)*n[0]*n[1], (gamma-1)*n[0]*n[2]],
                     [-gamma*beta*n[1], (gamma-1)*n[1]*n[0], 1+(gamma-1)*n[1]*n[1], (gamma-1)*n[1]*n[2]],
                     [-gamma*beta*n[2], (gamma-1)*n[2]*n[0], (gamma-1)*n[2]*n[1], 1+(gamma-1)*n[2]*n[2]]])


def lorentz_boost(momentum: LorentzVector, v: Vector3):
    """Lorentz boost `momentum` to a new frame with velocity v (reference: cross_section_mc.py:669-685).  Host helper for
    single 4-vectors; arrays of 4-vectors are boosted inside the kernels."""
    mat = _boost_matrix(v)
    if mat is None:
        return momentum
    boosted_p4 = mat @ momentum.pmu
    return LorentzVector(boosted_p4[0], boosted_p4[1], boosted_p4[2], boosted_p4[3])


class _Vec3Array:
    """N three-vectors as planes (v1, v2, v3): read-only sequence of `Vector3` (the reference keeps a list)."""

    def __init__(self, planes_device):
        self.device = planes_device
        self._host = None

    @property
    def vec(self):
        if self._host is None:
            self._host = planes_to_host_rows(self.device)
        return self._host

    def __len__(self):
        return int(self.device.shape[1])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [Vector3(*row) for row in self.vec[i]]
        return Vector3(*self.vec[i])

    def __iter__(self):
        for row in self.vec:
            yield Vector3(*row)

    def __array__(self, dtype=None, copy=None):
        return self.vec if dtype is None else self.vec.astype(dtype)


class Scatter2to2MC:
    """reference: cross_section_mc.py:148-275"""

    def __init__(self, mtrx2: MatrixElement2, p1=None, p2=None, n_samples=1000, seed=None, device=None):
        self.mtrx2 = mtrx2

        self.m1 = mtrx2.m1
        self.m2 = mtrx2.m2
        self.m3 = mtrx2.m3
        self.m4 = mtrx2.m4

        self.lv_p1 = LorentzVector() if p1 is None else p1
        self.lv_p2 = LorentzVector() if p2 is None else p2

        self.n_samples = n_samples
        self.seed = seed
        self._device = device
        self._calls = 0
        self._reset()

    def _reset(self):
        self.p3_cm_4vectors = []
        self.p3_lab_4vectors = []
        self.p3_cm_3vectors = []
        self.p3_lab_3vectors = []
        self.p4_lab_3vectors = []
        self.p4_lab_4vectors = []
        self.dsigma_dcos_cm_wgts = np.array([])
        self._dev = {}

    def set_new_scattter(self, new_p1: LorentzVector, new_p2: LorentzVector):
        self.lv_p1 = new_p1
        self.lv_p2 = new_p2
        self._reset()

    # ---- scalar kinematics: the reference's expressions --------------------------------------------------------------
    def _den(self, s):
        return 16*np.pi*(s - (self.m1 + self.m2)**2)*(s - (self.m1 - self.m2)**2)

    def dsigma_dt(self, s, t):
        """M2(s, t) / (16 pi (s - (m1+m2)^2)(s - (m1-m2)^2)) (cross_section_mc.py:184-185), evaluated on the device."""
        if np.ndim(s) == 0:
            return self.mtrx2._eval(s, t, 1.0, den=float(self._den(s)))
        return self.mtrx2._eval(s, t, 1.0) / self._den(np.asarray(s, dtype=np.float64))

    def p1_cm(self, s):
        return np.sqrt((np.power(s - self.m1**2 - self.m2**2, 2) - np.power(2*self.m1*self.m2, 2))/(4*s))

    def p3_cm(self, s):
        return np.sqrt((np.power(s - self.m3**2 - self.m4**2, 2) - np.power(2*self.m3*self.m4, 2))/(4*s))

    def _params(self, log_sampling=False):
        """Per-call constants of scatter_sim (alpk_scatter2_t), every scalar evaluated with the reference's expression
        (cross_section_mc.py:207-226); None below threshold."""
        # Declare momenta and energy in the CM frame (:207-213)
        cm_p4 = self.lv_p1 + self.lv_p2
        e_in = cm_p4.energy()
        p_in = cm_p4.get_3momentum()
        with np.errstate(all="ignore"):
            v_in = Vector3(p_in.v1 / e_in, p_in.v2 / e_in, p_in.v3 / e_in)
            s = cm_p4.mass2()
            if s < (self.m3 + self.m4)**2:
                return None

            p1_cm = self.p1_cm(s)
            p3_cm = self.p3_cm(s)
            e1_cm = np.sqrt(p1_cm**2 + self.m1**2)
            e3_cm = np.sqrt(p3_cm**2 + self.m3**2)

            par = N.Scatter2T()
            par.m2 = self.mtrx2._m2_struct(s)
            par.s = float(s)
            par.msum = float(self.m1**2 + self.m3**2)
            par.pp = float(p1_cm*p3_cm)
            par.ee = float(e1_cm*e3_cm)
            par.p3_cm, par.e3_cm = float(p3_cm), float(e3_cm)
            par.vol = float(4*p1_cm*p3_cm/self.n_samples)
            par.p1_cm, par.n_samples = float(p1_cm), float(self.n_samples)
            par.den = float(self._den(s))
            mat = _boost_matrix(-v_in)
        mat = np.eye(4) if mat is None else mat
        for k in range(16):
            par.mat[k] = float(mat.ravel()[k])
        for k in range(4):
            par.p12[k] = float(cm_p4.pmu[k])
        par.log_sampling = 1 if log_sampling else 0
        return par

    # ---- the Monte Carlo ----------------------------------------------------------------------------------------------
    def scatter_sim(self, log_sampling=False, uniforms=None):
        """Simulate n_samples scatterings p1 p2 -> p3 p4 (cross_section_mc.py:193-240).

        uniforms (parity mode): 2*n_samples U[0,1) variates -- the phi draws, then the theta (or logu) draws."""
        n = int(self.n_samples)
        self._reset()
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64).ravel()
            if uniforms.shape[0] != 2 * n:
                raise ValueError(f"uniforms must hold 2*n_samples = {2 * n} variates")
            seed = 0
        else:
            from .fluxes import _draw_seed
            seed = self.seed if self.seed is not None else _draw_seed()
        par = self._params(log_sampling)
        if par is None:                               # s < (m3 + m4)^2 (:212-213): nothing simulated
            return

        rt = Runtime.get(self._device)
        with rt.activate():
            d_u = rt.upload(uniforms) if (uniforms is not None and n) else None
            dev = {"t": rt.empty(n), "w": rt.empty(n), "cos": rt.empty(n)}
            for k in ("p3_cm", "p3_lab", "p4_lab"):
                dev[k] = torch.empty((4, n), dtype=torch.float64, device=rt.device)
            for k in ("v3_cm", "v3_lab", "v4_lab"):
                dev[k] = torch.empty((3, n), dtype=torch.float64, device=rt.device)
            if n:
                N.call("alpk_scatter2to2", C.byref(par), n, N.ptr(d_u), C.c_uint64(seed), self._calls & 0xFFFFFFFF,
                       N.ptr(dev["t"]), N.ptr(dev["w"]), N.ptr(dev["p3_cm"]), N.ptr(dev["p3_lab"]), N.ptr(dev["p4_lab"]),
                       N.ptr(dev["cos"]), N.ptr(dev["v3_cm"]), N.ptr(dev["v3_lab"]), N.ptr(dev["v4_lab"]), rt.stream())
        self._calls += 1
        self._dev = dev
        self.dsigma_dcos_cm_wgts = dev["w"].cpu().numpy()
        self.p3_cm_4vectors = LorentzVectorArray(dev["p3_cm"])
        self.p3_lab_4vectors = LorentzVectorArray(dev["p3_lab"])
        self.p4_lab_4vectors = LorentzVectorArray(dev["p4_lab"])
        self.p3_cm_3vectors = _Vec3Array(dev["v3_cm"])
        self.p3_lab_3vectors = _Vec3Array(dev["v3_lab"])
        self.p4_lab_3vectors = _Vec3Array(dev["v4_lab"])

    def device_arrays(self):
        """The last scatter_sim() on the device: t, w (weights), cos (lab cosine of p3), 4-vector planes p3_cm / p3_lab /
        p4_lab [4][n] and 3-velocity planes [3][n] as torch CUDA tensors (empty dict below threshold)."""
        return self._dev

    def get_total_xs(self, s):
        """Integral of dsigma/dt over the kinematic t range (cross_section_mc.py:242-258).  As in the reference the adaptive
        quadrature driver is scipy.integrate.quad on the host; every integrand evaluation is the device evaluator of the
        matrix element (dsigma_dt -> alpk_m2_eval), one small launch per abscissa -- a utility, not a hot-path call."""
        from numpy import power
        from scipy.integrate import quad
        p1_cm = self.p1_cm(s)
        p3_cm = self.p3_cm(s)

        t0 = power(self.m1**2 - self.m3**2 - self.m2**2 + self.m4**2, 2)/(4*s) - power(p1_cm - p3_cm, 2)
        t1 = power(self.m1**2 - self.m3**2 - self.m2**2 + self.m4**2, 2)/(4*s) - power(p1_cm + p3_cm, 2)

        def integrand(t):
            return float(self.dsigma_dt(s, t))

        return quad(integrand, t1, t0)[0]

    def get_cosine_lab_weights(self):
        """(cosines of p3 in the lab frame, weights) -- cross_section_mc.py:260-268"""
        if not self._dev:
            return np.array([]), self.dsigma_dcos_cm_wgts
        return self._dev["cos"].cpu().numpy(), self.dsigma_dcos_cm_wgts

    def get_e3_lab_weights(self):
        """(lab energies of p3, weights) -- cross_section_mc.py:270-272"""
        if not self._dev:
            return np.array([]), self.dsigma_dcos_cm_wgts
        return self._dev["p3_lab"][0].cpu().numpy(), self.dsigma_dcos_cm_wgts

    def get_e4_lab_weights(self):
        """(lab energies of p4, weights) -- cross_section_mc.py:274-276"""
        if not self._dev:
            return np.array([]), self.dsigma_dcos_cm_wgts
        return self._dev["p4_lab"][0].cpu().numpy(), self.dsigma_dcos_cm_wgts

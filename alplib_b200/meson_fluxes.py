"""
Dark-boson fluxes from charged-meson 3-body decays  M -> l nu X -- drop-in for the reference's
`FluxChargedMeson3BodyIsotropic` (fluxes.py:811-922) and `FluxChargedMeson3BodyDecay` (fluxes.py:605-808), with
`M2Meson3BodyDecay` (matrix_element.py:419-518).

Per sample the reference integrates the squared matrix element over the Dalitz variable m_23^2 with
`scipy.integrate.quad`; the kernels (csrc/alpk_meson.cu, csrc/alp_math3.cuh) evaluate the same 21-point Gauss-Kronrod
rule QUADPACK applies, including its acceptance test and bisection rule, one sample per thread.

Differences from the reference, all invisible to a script that reads the documented attributes after one simulate():
  * samples live on the GPU and are copied to the host lazily (as NumPy arrays instead of lists);
  * variates come from Philox keyed by (seed, row, sample); `simulate(uniforms=...)` injects U[0,1) variates (parity mode);
  * `FluxChargedMeson3BodyDecay.simulate()` resets `axion_angle` like the other per-sample lists (the reference keeps
    appending to it across calls); its `multicore=True` branch (a multiprocessing fan-out that raises TypeError in the
    reference whenever a meson decays beyond the dump) is not offered -- the GPU kernel is the parallel path.
There is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch
from numpy import arctan, cos, pi, power, sqrt

from . import _native as N
from .constants import (ALPHA, F_K, F_PI, KAON_WIDTH, M_E, M_ETA, M_ETA_PRIME, M_K, M_MU, M_PI, M_PI0, PION_WIDTH, V_UD,
                        V_US)
from .decay import W_agg_hadronic, W_ff, W_gg
from .fluxes import AxionFlux, _draw_seed, _split_kwargs
from .materials import Material

__all__ = ["FluxChargedMeson3BodyIsotropic", "FluxChargedMeson3BodyDecay", "M2Meson3BodyDecay", "FluxAxionMesonMixing"]

_MODELS = ["scalar_ib1", "pseudoscalar_ib1", "vector_ib1", "scalar_ib2", "vector_ib2", "vector_ib9", "sd", "vector_contact"]
_PARAMS = {"pion": [M_PI, V_UD, F_PI, PION_WIDTH], "kaon": [M_K, V_US, F_K, KAON_WIDTH]}      # fluxes.py:612-615


def _model_id(name):
    return _MODELS.index(name) if name in _MODELS else len(_MODELS)      # unknown string -> zero matrix element


def _meson3_params(mm, ma, fM, ckm, total_width, c0, coupling, abd, n_samples, model):
    p = N.Meson3T()
    p.mm, p.ma, p.fM, p.ckm, p.total_width = float(mm), float(ma), float(fM), float(ckm), float(total_width)
    p.c0, p.coupling = float(c0), float(coupling)
    p.a, p.b, p.d = (float(x) for x in abd)
    p.pref = float((2*mm)/(32*power(2*pi*mm, 3)))                        # fluxes.py:653
    p.n_samples = float(n_samples)
    p.model = _model_id(model)
    return p


class M2Meson3BodyDecay:
    """|M|^2 of M -> l(1) nu(2) X(3) (matrix_element.py:419-518); evaluated on the device."""

    def __init__(self, mx, meson="pion", lepton_mass=M_E, interaction_model="vector_ib2", abd=(0, 0, 0)):
        decay_params = _PARAMS[meson]
        self.m_parent = decay_params[0]
        self.abd = abd
        self.ckm = decay_params[1]
        self.fM = decay_params[2]
        self.total_width = decay_params[3]
        self.m1, self.m2, self.m3 = lepton_mass, 0.0, mx
        self.interaction = interaction_model

    def __call__(self, m122, m232, c0=0.0, coupling=1.0):
        from .runtime import Runtime
        rt = Runtime.get()
        x = np.atleast_1d(np.asarray(m232, dtype=np.float64))
        par = _meson3_params(self.m_parent, self.m3, self.fM, self.ckm, self.total_width, c0, coupling, self.abd, 1,
                             self.interaction)
        with rt.activate():
            d_x, out = rt.upload(x.ravel()), rt.empty(x.size)
            N.call("alpk_meson3_m2", float(m122), N.ptr(d_x), x.size, float(self.m1), C.byref(par), N.ptr(out), rt.stream())
        r = out.cpu().numpy().reshape(x.shape)
        return r if np.ndim(m232) else np.float64(r[0])


class _Meson3Flux(AxionFlux):
    """Shared device plumbing of the two charged-meson classes."""

    def _par(self):
        return _meson3_params(self.mm, self.ma, self.fM, self.ckm, self.total_width, self.c0, self.gmu,
                              getattr(self.m2_e, "abd", (0, 0, 0)), self.n_samples, self._model_name)

    def set_ma(self, ma):
        self.ma = ma
        self.m2_mu.m3 = self.ma
        self.m2_e.m3 = self.ma

    def dGammadEa(self, Ea, ml=M_MU):
        """dGamma/dE_a (fluxes.py:639-661 / 841-863): Dalitz integral of |M|^2, scalar or array E_a."""
        rt = self._rt
        ea = np.atleast_1d(np.asarray(Ea, dtype=np.float64))
        par = self._par()
        with rt.activate():
            d_e, out = rt.upload(ea.ravel()), rt.empty(ea.size)
            N.call("alpk_meson3_dgamma", N.ptr(d_e), ea.size, float(ml), C.byref(par), N.ptr(out), rt.stream())
        r = out.cpu().numpy().reshape(ea.shape)
        return r if np.ndim(Ea) else float(r[0])

    def total_br(self):
        """Total branching ratio (fluxes.py:663-667 / 865-866): outer adaptive quadrature on the host (scipy, as the
        reference), integrand = dGammadEa evaluated on the device."""
        from scipy.integrate import quad
        return np.sum([quad(self.dGammadEa, self.ma, (self.mm**2 + self.ma**2 - ml**2)/(2*self.mm), args=(ml,))[0]
                       for ml in self._br_leptons()]) / self.total_width

    def _rows(self, n_cols):
        """(meson table, row -> meson index, row -> lepton mass): one row per (meson, lepton with ma <= mm - ml), meson-major
        like the reference's loops (fluxes.py:743-749 / 906-912)."""
        mf = np.ascontiguousarray(np.asarray(self.meson_flux, dtype=np.float64))
        if mf.ndim != 2 or mf.shape[1] < n_cols:
            raise IndexError(f"meson_flux must have shape (n, {n_cols})")
        mf = np.ascontiguousarray(mf[:, :n_cols])
        leptons = [float(ml) for ml in self.lepton_masses if not (self.ma > self.mm - ml)]
        nm = mf.shape[0]
        row_meson = np.repeat(np.arange(nm, dtype=np.uint32), len(leptons))
        row_ml = np.tile(np.asarray(leptons, dtype=np.float64), nm)
        return mf, row_meson, row_ml


class FluxChargedMeson3BodyIsotropic(_Meson3Flux):
    """
    Isotropic dark-boson flux from stopped / slow charged mesons (reference: fluxes.py:811-922).  meson_flux rows are
    [momentum, weight]; per (meson, lepton) `n_samples` draws of (E_a, cos theta) in the meson rest frame, boosted along
    the meson direction.
    """
    def __init__(self, meson_flux=[[0.0, 0.0259]], boson_mass=0.1, coupling=1.0, meson_type="pion",
                 interaction_model="scalar_ib1", det_dist=20, det_length=2, det_area=2,
                 target=None, n_samples=50, c0=-0.97, lepton_masses=[M_E, M_MU], abd=(0, 0, 0), **kwargs):
        target = Material("W") if target is None else target
        super().__init__(boson_mass, target, det_dist, det_length, det_area, n_samples, **_split_kwargs(kwargs))
        self.meson_flux = meson_flux
        self.lepton_masses = lepton_masses
        decay_params = _PARAMS[meson_type]
        self.mm = decay_params[0]
        self.ckm = decay_params[1]
        self.fM = decay_params[2]
        self.total_width = decay_params[3]
        self.gmu = coupling
        self.n_samples = n_samples
        self.c0 = c0  # contact model parameter
        self._model_name = interaction_model
        self.m2_e = M2Meson3BodyDecay(boson_mass, meson_type, M_E, interaction_model, abd)
        self.m2_mu = M2Meson3BodyDecay(boson_mass, meson_type, M_MU, interaction_model, abd)
        self.row_offset = 0

    def _br_leptons(self):
        return self.lepton_masses

    def lifetime(self, gagamma):
        return 1/W_gg(gagamma, self.ma)

    def diff_br(self, uniforms=None):
        """(energies, weights) of `n_samples` rest-frame draws for `self.m_lepton` (fluxes.py:868-875; like the reference,
        the attribute must be set by the caller)."""
        ea_min = self.ma
        ea_max = (self.mm**2 + self.ma**2 - self.m_lepton**2)/(2*self.mm)
        mc_vol = ea_max - ea_min
        u = np.random.random_sample(int(self.n_samples)) if uniforms is None else np.asarray(uniforms, dtype=np.float64)
        energies = ea_min + (ea_max - ea_min)*u
        weights = mc_vol*self.dGammadEa(energies, self.m_lepton)/self.total_width/self.n_samples
        return energies, weights

    def simulate(self, uniforms=None):
        """uniforms: optional U[0,1) array in the reference's draw order -- per (meson, lepton) row `n_samples` energy
        variates then `n_samples` cosine variates (parity mode)."""
        self._reset_samples()
        mf, row_meson, row_ml = self._rows(2)
        n_rows, ns = int(row_meson.shape[0]), int(self.n_samples)
        rt = self._rt
        n = n_rows * ns
        seed = self.seed if self.seed is not None else (_draw_seed() if uniforms is None else 0)
        with rt.activate():
            E, F = rt.empty(n), rt.empty(n)
            if n:
                par = self._par()
                d_m, d_rm, d_ml = rt.upload(mf.ravel()), rt.upload(row_meson, np.uint32), rt.upload(row_ml)
                table = rt.empty(n_rows * N.MESON3_NP)
                N.call("alpk_meson3_rows", 0, N.ptr(d_m), int(mf.shape[0]), N.ptr(d_rm), N.ptr(d_ml), n_rows, C.byref(par),
                       float(self.det_dist), float(self.det_length), None, C.c_uint64(seed), 0, N.ptr(table), rt.stream())
                d_u = d_blk = None
                if uniforms is not None:
                    u = np.ascontiguousarray(uniforms, dtype=np.float64).ravel()
                    if u.shape[0] != 2 * n:
                        raise ValueError(f"uniforms must hold 2*n_rows*n_samples = {2 * n} variates, got {u.shape[0]}")
                    d_u, d_blk = rt.upload(u), rt.upload(np.arange(n_rows, dtype=np.int32), np.int32)
                N.call("alpk_meson3_samples", 0, N.ptr(table), n_rows, ns, C.byref(par), 0.0, 0, N.ptr(d_u), N.ptr(d_blk),
                       C.c_uint64(seed), int(self.row_offset), N.ptr(E), N.ptr(F), None, None, None, None, None, rt.stream())
        self._set_production(E, F)

    def propagate(self):
        """No decays; scatter weight = isotropic acceptance x flux, float64 like the reference (fluxes.py:917-920)."""
        self._ensure_production()
        rt = self._rt
        wgt = self._F if self._F is not None else torch.zeros(0, dtype=torch.float64, device=rt.device)
        n = int(wgt.shape[0])
        with rt.activate():
            d, s_ = rt.empty(n), rt.empty(n)
            N.call("alpk_scale", N.ptr(wgt), n, 0.0, N.ptr(d), rt.stream())
            N.call("alpk_scale", N.ptr(wgt), n, float(self.det_area / (4*pi*self.det_dist**2)), N.ptr(s_), rt.stream())
        self._D32, self._S32 = d, s_
        for k in ("D32", "S32"):
            self._host.pop(k, None)


class FluxChargedMeson3BodyDecay(_Meson3Flux):
    """
    Dark bosons from charged mesons decaying in flight between target and dump (reference: fluxes.py:605-808).
    meson_flux rows are [momentum, polar angle, weight]; every meson gets an exponential decay position, mesons decaying
    beyond 52 m produce nothing; per (meson, lepton) `n_samples` draws of (E_a, cos theta*, phi), kept when the boson
    points into the detector's solid angle (cut_on_solid_angle).
    """
    def __init__(self, meson_flux, boson_mass=0.1, coupling=1.0, n_samples=50, meson_type="pion",
                 interaction_model="scalar_ib1", energy_cut=140.0, det_dist=541, det_length=10.0,
                 det_area=25.0*pi, c0=-0.95, lepton_masses=[M_E, M_MU], verbose=False, **kwargs):
        super().__init__(boson_mass, Material("Be"), det_dist, det_length, det_area, n_samples=n_samples,
                         **_split_kwargs(kwargs))
        self.meson_flux = meson_flux
        decay_params = _PARAMS[meson_type]
        self.model = interaction_model
        self._model_name = interaction_model
        self.mm = decay_params[0]
        self.ckm = decay_params[1]
        self.fM = decay_params[2]
        self.total_width = decay_params[3]
        self.lepton_masses = lepton_masses
        self.gmu = coupling
        self.dump_dist = 50
        self.det_sa = cos(arctan(self.det_length/(self.det_dist-self.dump_dist)/2))     # shadows the method, as in the reference
        self.energy_cut = energy_cut
        self.c0 = c0
        self.m2_e = M2Meson3BodyDecay(boson_mass, meson_type, M_E, interaction_model)
        self.m2_mu = M2Meson3BodyDecay(boson_mass, meson_type, M_MU, interaction_model)
        self.verbose = verbose
        self.row_offset = 0
        self._extra = {}

    def _br_leptons(self):
        return [M_E, M_MU]                       # the reference's total_br is hard-wired to e and mu (fluxes.py:663-667)

    # per-sample lists of the reference, device resident, copied lazily
    def _extra_get(self, key):
        self._ensure_production()
        t = self._extra.get(key)
        if t is None:
            return []
        h = self._host.get(key)
        if h is None:
            h = self._host[key] = t.cpu().numpy()
        return h

    cosines = property(lambda self: self._extra_get("cosines"))
    solid_angles = property(lambda self: self._extra_get("solid_angles"))
    decay_pos = property(lambda self: self._extra_get("decay_pos"))

    @property
    def axion_angle(self):
        return self._extra_get("axion_angle")

    @axion_angle.setter
    def axion_angle(self, v):                    # AxionFlux.__init__ assigns []
        if isinstance(v, list) and not v:
            if hasattr(self, "_extra"):
                self._extra.pop("axion_angle", None)
            return
        self._extra["axion_angle"] = self._rt.upload(np.asarray(v, dtype=np.float64).ravel())
        self._host.pop("axion_angle", None)

    def simulate(self, cut_on_solid_angle=True, verbose=False, multicore=False, uniforms=None):
        """uniforms (parity mode): dict(pos=U[0,1) per meson, samples=U[0,1) blocks [live row][3][n_samples] in the
        reference's draw order: energies, cosines, azimuths)."""
        if multicore:
            raise NotImplementedError("multicore=True is a CPU fan-out of the reference; the CUDA kernel is the parallel path")
        self._reset_samples()
        self._extra = {}
        self._model_name = self.model
        mf, row_meson, row_ml = self._rows(3)
        n_rows, ns = int(row_meson.shape[0]), int(self.n_samples)
        rt = self._rt
        n = n_rows * ns
        seed = self.seed if self.seed is not None else (_draw_seed() if uniforms is None else 0)
        with rt.activate():
            if n == 0:
                z = rt.empty(0)
                self._set_production(z, z.clone())
                self._extra = {k: z.clone() for k in ("cosines", "axion_angle", "solid_angles", "decay_pos")}
                return
            par = self._par()
            d_m, d_rm, d_ml = rt.upload(mf.ravel()), rt.upload(row_meson, np.uint32), rt.upload(row_ml)
            table = rt.empty(n_rows * N.MESON3_NP)
            d_pos = None
            if uniforms is not None:
                up = np.ascontiguousarray(uniforms["pos"], dtype=np.float64).ravel()
                if up.shape[0] != mf.shape[0]:
                    raise ValueError("uniforms['pos'] must hold one variate per meson")
                d_pos = rt.upload(up)
            N.call("alpk_meson3_rows", 1, N.ptr(d_m), int(mf.shape[0]), N.ptr(d_rm), N.ptr(d_ml), n_rows, C.byref(par),
                   float(self.det_dist), float(self.det_length), N.ptr(d_pos), C.c_uint64(seed), int(self.row_offset),
                   N.ptr(table), rt.stream())
            d_u = d_blk = None
            if uniforms is not None:
                live = table.view(n_rows, N.MESON3_NP)[:, N.MESON3_NP - 1].cpu().numpy() != 0.0
                blk = np.where(live, np.cumsum(live) - 1, -1).astype(np.int32)
                us = np.ascontiguousarray(uniforms["samples"], dtype=np.float64).ravel()
                if us.shape[0] != 3 * ns * int(live.sum()):
                    raise ValueError(f"uniforms['samples'] must hold 3*n_samples per live row = {3 * ns * int(live.sum())} variates")
                d_u, d_blk = rt.upload(us), rt.upload(blk, np.int32)
            E, F, Cc, Th, Sa, Po = (rt.empty(n) for _ in range(6))
            keep = torch.zeros(n, dtype=torch.uint8, device=rt.device)
            N.call("alpk_meson3_samples", 1, N.ptr(table), n_rows, ns, C.byref(par), float(self.energy_cut),
                   1 if cut_on_solid_angle else 0, N.ptr(d_u), N.ptr(d_blk), C.c_uint64(seed), int(self.row_offset),
                   N.ptr(E), N.ptr(F), N.ptr(Cc), N.ptr(Th), N.ptr(Sa), N.ptr(Po), N.ptr(keep), rt.stream())
            sel = torch.nonzero(keep, as_tuple=False).ravel()
            E, F, Cc, Th, Sa, Po = (t.index_select(0, sel).contiguous() for t in (E, F, Cc, Th, Sa, Po))
        self._set_production(E, F)
        self._extra = {"cosines": Cc, "axion_angle": Th, "solid_angles": Sa, "decay_pos": Po}

    def propagate(self, decay=None, new_coupling=None):
        """decay truthy ('photon' / 'electron' in the reference's comment): survival x decay weights with the summed
        a -> l l widths (fluxes.py:792-802); otherwise no decays, scatter weight = flux in float64 (:803-806)."""
        if decay:
            if new_coupling is not None:
                width = np.sum([W_ff(new_coupling, mf, self.ma) for mf in self.lepton_masses])
                rescale = power(new_coupling/self.gmu, 2)
                super().propagate(width, rescale)
            else:
                width = np.sum([W_ff(self.gmu, mf, self.ma) for mf in self.lepton_masses])
                super().propagate(width)
        else:
            self._ensure_production()
            rt = self._rt
            wgt = self._F if self._F is not None else torch.zeros(0, dtype=torch.float64, device=rt.device)
            n = int(wgt.shape[0])
            with rt.activate():
                d, s_ = rt.empty(n), rt.empty(n)
                N.call("alpk_scale", N.ptr(wgt), n, 0.0, N.ptr(d), rt.stream())
                N.call("alpk_scale", N.ptr(wgt), n, 1.0, N.ptr(s_), rt.stream())
            self._D32, self._S32 = d, s_
            for k in ("D32", "S32"):
                self._host.pop(k, None)


class FluxAxionMesonMixing(AxionFlux):
    """
    ALP flux from pi0 / eta mixing (reference: fluxes.py:1148-1240, after Kelly, Kumar, Liu 2011.05995); f_a in MeV.
    meson_flux rows are 4-vectors [E, px, py, pz]; a meson above the ALP mass and inside the detector's half opening angle
    becomes one ALP of the same energy with the mixing-suppressed rate.
    """
    def __init__(self, meson_flux=None, meson_mass=M_PI0, target=None,
                 det_dist=574.0, det_length=5.0, det_area=21.0,
                 axion_mass=0.1, f_a=1000.0, meson_type="Pi0", total_pot=1e21,
                 n_samples=1, mesons_per_pot=2.89, c1=1.0, c2=1.0, c3=1.0, **kwargs):
        self.c1 = c1
        self.c2 = c2
        self.c3 = c3
        self.f_a = f_a
        target = Material("C") if target is None else target
        super().__init__(axion_mass, target, det_dist, det_length, det_area, **_split_kwargs(kwargs))
        self.axion_coupling = self.photon_coupling(f_a)
        self.meson_flux = meson_flux
        self.n_samples = n_samples
        self.support = np.ones(n_samples)
        self.meson_mass = meson_mass
        self.mesons_per_pot = mesons_per_pot
        self.total_pot = total_pot
        self.meson_norm_wgt = 1.0/meson_flux.shape[0]
        self.meson_type = meson_type
        self._angle = None

    def photon_coupling(self, f_a):
        # ma < QCD scale
        c_gamma = self.c2 + (5/3)*self.c1 + self.c3 * (
            -1.92
            + (1/3) * (self.ma**2 / (self.ma**2 - M_PI0**2))
            + (8/9) * ((self.ma**2 - (4/9)*M_PI0**2) / (self.ma**2 - M_ETA**2))
            + (7/9) * ((self.ma**2 - (16/9)*M_PI0**2) / (self.ma**2 - M_ETA_PRIME**2))
        )
        return c_gamma * ALPHA / (2 * pi * f_a)  # in MeV^-1

    def set_new_coupling(self, f_a):
        self.f_a = f_a
        self.axion_coupling = self.photon_coupling(f_a)

    def mixing(self):
        if self.meson_type == "Pi0":
            return (1/6) * (F_PI/self.f_a) * (self.ma**2 / (self.ma**2 - M_PI0**2))
        elif self.meson_type == "Eta":
            return (1/sqrt(6)) * (F_PI/self.f_a) * ((self.ma**2 - 4*M_PI0**2 / 9) / (self.ma**2 - M_ETA**2))

    def phase_space_factor(self):
        if self.ma <= self.meson_mass:
            return 1.0
        return np.power(self.ma / self.meson_mass, -1.6)

    @property
    def axion_angle(self):
        self._ensure_production()
        t = getattr(self, "_angle", None)
        if t is None:
            return []
        h = self._host.get("angle")
        if h is None:
            h = self._host["angle"] = t.cpu().numpy()
        return h

    @axion_angle.setter
    def axion_angle(self, v):
        if isinstance(v, list) and not v:
            self._angle = None
            return
        self._angle = self._rt.upload(np.asarray(v, dtype=np.float64).ravel())
        if hasattr(self, "_host"):
            self._host.pop("angle", None)

    def simulate(self):
        self._reset_samples()
        self._angle = None
        mf = np.ascontiguousarray(np.asarray(self.meson_flux, dtype=np.float64))
        if mf.ndim != 2 or mf.shape[1] < 4:
            raise IndexError("meson_flux must have shape (n, 4): rows of [E, px, py, pz]")
        mf = np.ascontiguousarray(mf[:, :4])
        n = int(mf.shape[0])
        rate = self.total_pot * self.mesons_per_pot * self.meson_norm_wgt \
            * self.phase_space_factor() * (self.mixing())**2                         # fluxes.py:1222-1223
        rt = self._rt
        with rt.activate():
            E, Th, F = rt.empty(n), rt.empty(n), rt.empty(n)
            if n:
                keep = torch.zeros(n, dtype=torch.uint8, device=rt.device)
                d_m = rt.upload(mf.ravel())
                N.call("alpk_meson_mixing", N.ptr(d_m), n, float(self.ma), float(self.det_sa()), float(rate), N.ptr(E),
                       N.ptr(Th), N.ptr(F), N.ptr(keep), rt.stream())
                sel = torch.nonzero(keep, as_tuple=False).ravel()
                E, Th, F = (t.index_select(0, sel).contiguous() for t in (E, Th, F))
        self._set_production(E, F)
        self._angle = Th

    def propagate(self, new_fa=None):
        """fluxes.py:1226-1236 (without new_fa the reference propagates with the photon width only)."""
        if new_fa is not None:
            total_width = W_gg(self.photon_coupling(new_fa), self.ma) + W_agg_hadronic(new_fa, self.ma)
            rescale = power(self.f_a/new_fa, 2)
            super().propagate(total_width, rescale)
        else:
            super().propagate(W_gg(self.axion_coupling, self.ma))

"""
ALP flux generators -- drop-in for the reference's `fluxes.py` classes on the Monte-Carlo hot path.

Same constructor signatures, attributes and method names as the reference (cited per class).  Differences, all
opt-in or invisible to a script that only uses the documented attributes:

  * samples live on the GPU; `axion_energy`, `axion_flux` (float64) and `decay_axion_weight`, `scatter_axion_weight`
    (float32, like the reference) are materialised on the host lazily, as NumPy arrays, when read.  Assigning to
    them uploads the new values.
  * random variates come from a counter-based Philox4x32-10 stream keyed by (seed, row, sample) instead of the global
    MT19937 state; `seed=None` draws the seed from `np.random` so `np.random.seed(...)` still makes runs
    reproducible.  `simulate(uniforms=...)` injects U[0,1) variates in the reference's draw order (parity mode).
  * `precision="fast"` (default, `alplib_b200.config.PRECISION`) evaluates the same quantities with algebraically
    simplified arithmetic (one reciprocal square root instead of seven divisions in `propagate`, a shared reciprocal in
    the Compton spectrum); results agree with the reference to <= 1e-9 relative.  `precision="faithful"` follows the
    reference's operation order expression by expression (bit-identical wherever libm allows).  `precision="fp32"` is the
    optional single-precision mode of the north star: production stays FP64, the survival / decay probabilities use
    float rsqrt / exp / expm1 (weights within ~1e-6 of the reference where its own `1 - exp(-x)` is well conditioned).
  * execution is deferred by default (`fused="auto"`, `alplib_b200.config.FUSED`): `simulate()` / `propagate()` record
    their arguments and kernels are launched when a result is demanded -- `decays()` then runs production -> propagate
    -> decay weights in ONE kernel that also materialises every per-sample array in HBM; reading `axion_energy` etc.
    earlier launches just what that attribute needs.  The variates are a pure function of the seed, so every schedule
    gives bit-identical arrays.  `fused=False` launches one kernel per call (the reference's schedule); `fused=True`
    never materialises per-sample arrays unless they are read (reduce-only decay rates).

There is no CPU fallback: every array operation below is a CUDA kernel in libalpk.so.
"""
import ctypes as C

import numpy as np
import torch
from numpy import arctan, pi, power, sqrt

from . import _native as N
from . import config
from . import params as P
from .constants import ALPHA, AVOGADRO, HBARC, M_E, METER_BY_MEV
from .decay import W_ee, W_gg, W_gg_loop
from .materials import DetectorGeometry, Material
from .photon_xs import AbsCrossSection
from .runtime import Runtime

_EMPTY64 = np.zeros(0, dtype=np.float64)

__all__ = ["AxionFlux", "FluxPrimakoffIsotropic", "FluxComptonIsotropic", "FluxBremIsotropic",
           "FluxResonanceIsotropic", "FluxPairAnnihilationIsotropic", "FluxNuclearIsotropic", "FluxPi0Isotropic",
           "FluxPairAnnihilationGamma", "FluxNeutralMeson2BodyDecay", "track_length_prob"]


def _as_flux_table(flux, name):
    a = np.asarray(flux, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] < 2:
        # the reference indexes photon[0], photon[1] of every row (fluxes.py:134-135) and fails on 1-D input
        raise IndexError(f"{name} must have shape (n, 2): rows of [energy, weight]")
    return a


def _kept_columns(pf, keep):
    """(row indices, energy column, weight column) of the rows with keep == True, as cheap views where possible:
    a threshold cut on a sorted spectrum keeps one contiguous block, which is sliced instead of gathered."""
    idx = np.flatnonzero(keep)
    n = int(idx.shape[0])
    if n and int(idx[-1]) - int(idx[0]) + 1 == n:
        sl = slice(int(idx[0]), int(idx[-1]) + 1)
        return idx, pf[sl, 0], pf[sl, 1]
    return idx, np.take(pf[:, 0], idx), np.take(pf[:, 1], idx)


def _draw_seed():
    # one draw from the global legacy NumPy RNG, so np.random.seed(s) pins the Philox key like it pins the reference
    return int(np.random.randint(0, 2**63 - 1, dtype=np.int64))


class AxionFlux:
    """Generic superclass (reference: fluxes.py:20-89)."""

    def __init__(self, axion_mass, target: Material, det_dist, det_length, det_area, n_samples=1000,
                 off_axis_angle=0.0, timing_window_ns=np.inf, seed=None, device=None, fused=None, keep_f64=False,
                 precision=None, strict=None):
        self.ma = axion_mass
        self.target_z = target.z[0]
        self.target_a = target.z[0] + target.n[0]
        self.target_density = target.density
        self.det_dist = det_dist  # meters
        self.det_length = det_length  # meters
        self.det_area = det_area  # square meters
        self.axion_angle = []
        self.n_samples = n_samples
        self.off_axis_angle = off_axis_angle
        self.timing_window = 1e-9 * timing_window_ns
        # --- device state ---
        self.seed = seed
        self.fused = config.FUSED if fused is None else fused      # False | True | "auto" (see config.py)
        self.keep_f64 = keep_f64          # also keep the float64 weights before the float32 cast (parity / fp64 mode)
        # strict compatibility (config.STRICT): axion_energy / axion_flux are materialised as Python lists of np.float64
        # (what the reference's simulate() leaves behind, fluxes.py:222-224), so `.append`, `+ [..]` etc. behave as there
        self.strict = config.STRICT if strict is None else bool(strict)
        self.precision = precision if precision is not None else config.PRECISION
        if self.precision not in ("fast", "faithful", "fp32"):
            raise ValueError("precision must be 'fast', 'faithful' or 'fp32'")
        self.rng_rounds = config.PHILOX_ROUNDS[self.precision]     # Philox rounds of the production sample stream (config.py)
        self._device = device
        self._rt_cached = None
        self._reset_samples()

    # ---- runtime / state -----------------------------------------------------------------------------------------
    @property
    def _rt(self) -> Runtime:
        if self._rt_cached is None:
            self._rt_cached = Runtime.get(self._device)
        return self._rt_cached

    _epoch = 0                              # bumped by every simulate() / propagate(): lets event generators detect stale state

    def _reset_samples(self):
        self._epoch += 1
        self._E = self._F = None            # device float64: axion_energy / axion_flux
        self._D32 = self._S32 = None        # device float32: decay / scatter weights
        self._D64 = self._S64 = None        # device float64 pre-cast weights (keep_f64)
        self._pending_production = None     # True: a deferred simulate() still has to run self._produce() (fused mode); a flag,
                                            # not the bound method -- that would make every flux object a reference cycle
                                            # whose GBs of device arrays wait for the cyclic garbage collector
        self._pending_prop = None           # (PropT) recorded by propagate() in fused mode
        self._host = {}

    def n_flux(self):
        self._ensure_production()
        return 0 if self._E is None else int(self._E.shape[0])

    def _ensure_production(self):
        if self._E is None and self._pending_production is not None:
            self._produce()

    def _ensure_propagated(self):
        if self._D32 is None and self._pending_prop is not None:
            self._run_propagate(self._pending_prop)

    def _set_production(self, E, F):
        self._E, self._F = E, F
        self._D32 = self._S32 = self._D64 = self._S64 = None
        self._host = {}

    def _host_get(self, key, tensor, empty):
        if tensor is None:
            return empty
        h = self._host.get(key)
        if h is None:
            # pinned, chunked, double-buffered device -> host copy (Runtime.to_host)
            h = self._rt.to_host(tensor)
            if self.strict and key in ("E", "F"):
                # the reference's simulate() leaves Python lists of np.float64 (fluxes.py:222-224: list.extend / append)
                h = list(h)
            self._host[key] = h
        return h

    # ---- the reference's public attributes ---------------------------------------------------------------------------
    @property
    def axion_energy(self):
        self._ensure_production()
        return self._host_get("E", self._E, [])

    @axion_energy.setter
    def axion_energy(self, v):
        self._ensure_production()            # a deferred simulate() must not overwrite (or lose) the other array
        self._pending_production = None
        self._E = self._rt.upload(np.asarray(v, dtype=np.float64).ravel())
        self._host.pop("E", None)

    @property
    def axion_flux(self):
        self._ensure_production()
        return self._host_get("F", self._F, [])

    @axion_flux.setter
    def axion_flux(self, v):
        self._ensure_production()
        self._pending_production = None
        self._F = self._rt.upload(np.asarray(v, dtype=np.float64).ravel())
        self._host.pop("F", None)

    @property
    def decay_axion_weight(self):
        self._ensure_propagated()
        return self._host_get("D32", self._D32, [])

    @decay_axion_weight.setter
    def decay_axion_weight(self, v):
        self._ensure_production()
        self._ensure_propagated()
        self._pending_prop = None
        self._D32 = self._rt.upload(np.asarray(v, dtype=np.float32).ravel(), np.float32)
        self._host.pop("D32", None)

    @property
    def scatter_axion_weight(self):
        self._ensure_propagated()
        return self._host_get("S32", self._S32, [])

    @scatter_axion_weight.setter
    def scatter_axion_weight(self, v):
        self._ensure_production()
        self._ensure_propagated()
        self._pending_prop = None
        self._S32 = self._rt.upload(np.asarray(v, dtype=np.float32).ravel(), np.float32)
        self._host.pop("S32", None)

    @property
    def decay_axion_weight_f64(self):
        """float64 decay weights before the float32 cast and acceptance factor (needs keep_f64=True)."""
        self._ensure_propagated()
        return self._host_get("D64", self._D64, _EMPTY64)

    @property
    def scatter_axion_weight_f64(self):
        self._ensure_propagated()
        return self._host_get("S64", self._S64, _EMPTY64)

    def device_arrays(self):
        """The sample arrays as torch CUDA tensors (no copy): E, flux (f64), decay_w, scatter_w (f32)."""
        self._ensure_production()
        self._ensure_propagated()
        return {"axion_energy": self._E, "axion_flux": self._F, "decay_axion_weight": self._D32,
                "scatter_axion_weight": self._S32}

    # ---- reference methods ----------------------------------------------------------------------------------------------
    def det_sa(self):
        return arctan(sqrt(self.det_area / pi) / self.det_dist)

    def _prop_params(self, decay_width, rescale_factor, geom_accept, time_of_flight_cut):
        return P.prop_params(self.ma, decay_width, rescale_factor, self.det_dist, self.det_length,
                             geom_accept=geom_accept,
                             timing_window=self.timing_window if time_of_flight_cut is not None else None,
                             fast="fp32" if self.precision == "fp32" else self.precision == "fast")

    def _run_propagate(self, prop):
        self._ensure_production()
        self._pending_prop = None
        rt = self._rt
        n = 0 if self._E is None else int(self._E.shape[0])
        if (0 if self._F is None else int(self._F.shape[0])) != n:
            raise ValueError("axion_energy and axion_flux have different lengths")
        with rt.activate():
            self._D32 = rt.empty(n, torch.float32)
            self._S32 = rt.empty(n, torch.float32)
            self._D64 = rt.empty(n) if self.keep_f64 else None
            self._S64 = rt.empty(n) if self.keep_f64 else None
            if n:
                N.call("alpk_propagate", N.ptr(self._E), N.ptr(self._F), n, C.byref(prop), N.ptr(self._D32),
                       N.ptr(self._S32), N.ptr(self._D64), N.ptr(self._S64), rt.stream())
        for k in ("D32", "S32", "D64", "S64"):
            self._host.pop(k, None)

    def _propagate(self, decay_width, rescale_factor=1.0, geom_accept=None, time_of_flight_cut=None):
        self._epoch += 1
        prop = self._prop_params(decay_width, rescale_factor, geom_accept, time_of_flight_cut)
        self._D32 = self._S32 = self._D64 = self._S64 = None
        for k in ("D32", "S32", "D64", "S64"):
            self._host.pop(k, None)
        if self.fused:
            self._pending_prop = prop
        else:
            self._run_propagate(prop)

    def propagate(self, decay_width, rescale_factor=1.0, time_of_flight_cut=None):
        """Survival x decay probability weights (reference: AxionFlux.propagate, fluxes.py:43-65)."""
        self._propagate(decay_width, rescale_factor, None, time_of_flight_cut)

    def propagate_iso_vol_int(self, geom: DetectorGeometry, decay_width, rescale_factor=1.0):
        """Decay weights from a Monte-Carlo integral of the decay probability density over the detector volume
        (reference: AxionFlux.propagate_iso_vol_int, fluxes.py:67-89): n_flux x geom.n_samples evaluations."""
        self._ensure_production()
        self._pending_prop = None
        rt = self._rt
        n = 0 if self._E is None else int(self._E.shape[0])
        if (0 if self._F is None else int(self._F.shape[0])) != n:
            raise ValueError("axion_energy and axion_flux have different lengths")
        prop = self._prop_params(decay_width, rescale_factor, None, None)
        ls, jac, vol = geom.point_table()
        npts = int(ls.shape[0])
        with rt.activate():
            self._D32 = rt.empty(n, torch.float32)
            self._S32 = rt.empty(n, torch.float32)
            self._D64 = rt.empty(n) if self.keep_f64 else None
            self._S64 = rt.empty(n) if self.keep_f64 else None
            if n:
                d_nx = rt.upload(-ls / METER_BY_MEV)                     # fluxes.py:85  -x / METER_BY_MEV
                d_b = rt.upload(4*pi*ls**2)                              # :85  4*pi*x**2
                d_j = rt.upload(jac)
                chunks = int(N.lib().alpk_vol_int_chunks(n, npts))
                partials = rt.empty(chunks * n)
                N.call("alpk_vol_int", N.ptr(self._E), N.ptr(self._F), n, C.byref(prop), N.ptr(d_nx), N.ptr(d_b),
                       N.ptr(d_j), npts, float(vol / geom.n_samples), float(1 / METER_BY_MEV), N.ptr(self._D32),
                       N.ptr(self._S32), N.ptr(self._D64), N.ptr(self._S64), N.ptr(partials), rt.stream())
        for k in ("D32", "S32", "D64", "S64"):
            self._host.pop(k, None)

    def _geom_accept(self):
        return self.det_area / (4*pi*self.det_dist**2)          # fluxes.py:160

    # ---- event-rate hooks used by the generators -------------------------------------------------------------------------
    def _decay_rate(self, ev, want_event_w):
        """(rate tensor[1], event_w tensor or None): generators.py:88-92 on the device."""
        rt = self._rt
        if self.fused is True and not want_event_w:
            fused = self._fused_decay_rate(ev)
            if fused is not None:
                return fused, None
        elif self.fused == "auto":
            both = self._fused_decay_materialize(ev, want_event_w)
            if both is not None:
                return both
        self._ensure_production()
        self._ensure_propagated()
        n = 0 if self._E is None else int(self._E.shape[0])
        if self._D32 is None or int(self._D32.shape[0]) != n:
            raise ValueError("propagate() must run before decays(): decay_axion_weight does not match axion_energy")
        with rt.activate():
            rate = rt.empty(1)
            evw = rt.empty(n) if want_event_w else None
            N.call("alpk_decays", N.ptr(self._E), N.ptr(self._D32), n, C.byref(ev), N.ptr(evw), N.ptr(rate),
                   N.ptr(rt.scratch), rt.stream())
        return rate, evw

    def _fused_decay_rate(self, ev):
        return None      # flux classes with a fused production->propagate->decays kernel override this

    def _fused_decay_materialize(self, ev, want_event_w):
        return None


def _width_e(self, ge, ma):
    if self.loop_decay:
        return W_ee(ge, ma) + W_gg_loop(ge, ma, M_E)
    else:
        return W_ee(ge, ma)


class FluxPrimakoffIsotropic(AxionFlux):
    """
    Primakoff-produced ALP flux from a photon flux table [[E_gamma, counts/s], ...] in the target
    (reference: fluxes.py:111-169).  One ALP sample per photon row with E_gamma >= m_a.
    """
    def __init__(self, photon_flux=[[1, 1]], target=None, det_dist=4.0, det_length=0.2,
                 det_area=0.04, axion_mass=0.1, axion_coupling=1e-3, n_samples=1000, **kwargs):
        target = Material("W") if target is None else target
        super().__init__(axion_mass, target, det_dist, det_length, det_area, **kwargs)
        self.photon_flux = photon_flux
        self.gagamma = axion_coupling
        self.n_samples = n_samples
        self.target_photon_xs = AbsCrossSection(target)

    _TABLE_ATTR = "photon_flux"

    def _row_keep(self, table):
        """Rows that produce a sample: fluxes.py:136 returns for E_gamma < m_a."""
        return ~(table[:, 0] < self.ma)

    def decay_width(self, gagamma, ma):
        return W_gg(gagamma, ma)

    def photon_flux_dN_dE(self, energy):
        pf = np.asarray(self.photon_flux, dtype=np.float64)
        return np.interp(energy, pf[:, 0], pf[:, 1], left=0.0, right=0.0)

    def _stage_inputs(self):
        """Host-side row filter (fluxes.py:136) + upload of the surviving rows (SoA: energies, weights)."""
        pf = _as_flux_table(self.photon_flux, "photon_flux")
        rt = self._rt
        with rt.activate():
            if pf.flags.c_contiguous:
                d_e, d_w, _, n = rt.stage_flux_rows(pf, 0, float(self.ma))
            else:
                keep = ~(pf[:, 0] < self.ma)
                idx, e_k, w_k = _kept_columns(pf, keep)
                d_e, d_w = rt.upload_packed([(e_k, np.float64), (w_k, np.float64)])
                n = int(idx.shape[0])
            self._inputs = (d_e, d_w, n)
        return self._inputs

    def _produce(self):
        """Launch the per-row production kernel on the staged inputs (device resident)."""
        self._pending_production = None
        d_eg, d_ph, n = self._inputs
        rt = self._rt
        with rt.activate():
            E, F = rt.empty(n), rt.empty(n)
            if n:
                le, lx, nt = self.target_photon_xs.device_table(rt)
                pref, r02 = P.primakoff_consts(self.gagamma, self.target_z)
                N.call("alpk_primakoff", N.ptr(d_eg), N.ptr(d_ph), n, N.ptr(le), N.ptr(lx), nt,
                       float(self.ma), pref, r02, N.ptr(E), N.ptr(F), rt.stream())
        self._set_production(E, F)

    def simulate(self):
        self._reset_samples()
        self._stage_inputs()
        if self.fused:
            self._pending_production = True
        else:
            self._produce()

    # ---- one-kernel schedule: production -> propagate -> decay weights + rate (alpk_primakoff_chain) ----------------------------
    def _launch_chain(self, prop, ev, materialize, want_event_w):
        d_eg, d_ph, n = self._inputs
        rt = self._rt
        with rt.activate():
            rate = rt.empty(1)
            E = F = d32 = s32 = d64 = s64 = evw = None
            if materialize:
                E, F = rt.empty(n), rt.empty(n)
                d32, s32 = rt.empty(n, torch.float32), rt.empty(n, torch.float32)
                if self.keep_f64:
                    d64, s64 = rt.empty(n), rt.empty(n)
            if want_event_w:
                evw = rt.empty(n)
            le, lx, nt = self.target_photon_xs.device_table(rt)
            pref, r02 = P.primakoff_consts(self.gagamma, self.target_z)
            N.call("alpk_primakoff_chain", N.ptr(d_eg), N.ptr(d_ph), n, N.ptr(le), N.ptr(lx), nt, float(self.ma), pref, r02,
                   C.byref(prop), C.byref(ev), N.ptr(E), N.ptr(F), N.ptr(d32), N.ptr(s32), N.ptr(d64), N.ptr(s64), N.ptr(evw),
                   N.ptr(rate), N.ptr(rt.scratch), rt.stream())
        if materialize:
            self._pending_production = None
            self._pending_prop = None
            self._set_production(E, F)
            self._D32, self._S32, self._D64, self._S64 = d32, s32, d64, s64
        return rate, evw

    def _chain_pending(self):
        return (getattr(self, "_inputs", None) is not None and self._pending_production is not None
                and self._pending_prop is not None and self._E is None)

    def _fused_decay_rate(self, ev):
        """fused=True: the rate alone, nothing materialised."""
        if not self._chain_pending():
            return None
        return self._launch_chain(self._pending_prop, ev, False, False)[0]

    def _fused_decay_materialize(self, ev, want_event_w):
        """"auto" schedule: one kernel that also writes every per-sample array."""
        if not self._chain_pending():
            return None
        return self._launch_chain(self._pending_prop, ev, True, want_event_w)

    def chain(self, days_exposure, threshold, new_coupling=None, materialize=True):
        """One-kernel production -> propagate -> decays on the staged photon rows (simulate() first): returns (rate tensor,
        event-weight tensor or None); with materialize=True also fills axion_energy / axion_flux / decay_axion_weight /
        scatter_axion_weight on the device."""
        if getattr(self, "_inputs", None) is None:
            raise RuntimeError("simulate() must be called first")
        g = self.gagamma if new_coupling is None else new_coupling
        rescale = 1.0 if new_coupling is None else power(new_coupling/self.gagamma, 2)
        prop = self._prop_params(W_gg(g, self.ma), rescale, self._geom_accept(), None)
        return self._launch_chain(prop, P.event_params(days_exposure, threshold), materialize, materialize)

    def propagate(self, new_coupling=None, is_isotropic=True):
        geom = self._geom_accept() if is_isotropic else None
        if new_coupling is not None:
            rescale = power(new_coupling/self.gagamma, 2)
            self._propagate(W_gg(new_coupling, self.ma), rescale, geom)
        else:
            self._propagate(W_gg(self.gagamma, self.ma), 1.0, geom)

    def propagate_iso_vol_int(self, geom: DetectorGeometry, new_coupling=None):
        if new_coupling is not None:
            rescale = power(new_coupling/self.gagamma, 2)
            super().propagate_iso_vol_int(geom, W_gg(new_coupling, self.ma), rescale)
        else:
            super().propagate_iso_vol_int(geom, W_gg(self.gagamma, self.ma))


def _vol_int_e(self, geom: DetectorGeometry, new_coupling=None):
    """propagate_iso_vol_int of the electron-coupled flux classes (fluxes.py:247-252, 370-375, 459-464, 544-549):
    uses W_ee only, like the reference (no loop-induced width here)."""
    if new_coupling is not None:
        rescale = power(new_coupling/self.ge, 2)
        AxionFlux.propagate_iso_vol_int(self, geom, W_ee(new_coupling, self.ma), rescale)
    else:
        AxionFlux.propagate_iso_vol_int(self, geom, W_ee(self.ge, self.ma))


class _ChainFlux(AxionFlux):
    """Flux classes whose samples come from the generic per-sample chain kernel (production -> [propagate -> [decays]]):
    Compton, bremsstrahlung, associated production.  Subclasses stage their inputs and build the per-row table."""
    _SAMPLES_FN = None          # C entry point of the per-sample kernel
    _rows = None                # (row table, n_rows, n_samples, row ids, params struct, injected uniforms, seed)

    decay_width = _width_e
    propagate_iso_vol_int = _vol_int_e

    def _launch_samples(self, E, F, prop, ev, d32=None, s32=None, d64=None, s64=None, evw=None, rate=None,
                        hist_edges=None, hist=None):
        table, n_rows, ns, d_ids, par, d_u, seed = self._rows
        rt = self._rt
        N.call(self._SAMPLES_FN, N.ptr(table), n_rows, ns, N.ptr(d_ids), C.byref(par), N.ptr(d_u),
               C.c_uint64(seed), N.ptr(E), N.ptr(F),
               C.byref(prop) if prop is not None else None, N.ptr(d32), N.ptr(s32), N.ptr(d64), N.ptr(s64),
               C.byref(ev) if ev is not None else None, N.ptr(evw), N.ptr(rate), N.ptr(rt.scratch),
               N.ptr(hist_edges), 0 if hist_edges is None else int(hist_edges.shape[0]), N.ptr(hist), rt.stream())

    def _produce(self):
        self._pending_production = None
        rt = self._rt
        n = self._rows[1] * self._rows[2]
        with rt.activate():
            E, F = rt.empty(n), rt.empty(n)
            self._launch_samples(E, F, None, None)
        self._set_production(E, F)

    def _finish_simulate(self):
        if self.fused:
            self._pending_production = True
        else:
            self._produce()

    def _fused_decay_rate(self, ev):
        """production -> propagate -> decays in one kernel, nothing materialised."""
        if self._rows is None or self._pending_prop is None or self._E is not None:
            return None
        rt = self._rt
        with rt.activate():
            rate = rt.empty(1)
            self._launch_samples(None, None, self._pending_prop, ev, rate=rate)
        return rate

    def _fused_decay_materialize(self, ev, want_event_w):
        """"auto" schedule: production -> propagate -> decays in one kernel that also writes every per-sample array."""
        if self._rows is None or self._pending_prop is None or self._E is not None or self.keep_f64:
            return None
        rt = self._rt
        n = self._rows[1] * self._rows[2]
        prop = self._pending_prop
        with rt.activate():
            rate = rt.empty(1)
            E, F = rt.empty(n), rt.empty(n)
            d32, s32 = rt.empty(n, torch.float32), rt.empty(n, torch.float32)
            evw = rt.empty(n) if want_event_w else None
            self._launch_samples(E, F, prop, ev, d32=d32, s32=s32, evw=evw, rate=rate)
        self._pending_production = None
        self._pending_prop = None
        self._set_production(E, F)
        self._D32, self._S32 = d32, s32
        return rate, evw

    def chain(self, days_exposure, threshold, new_coupling=None, materialize=True, hist_edges=None, hist=None, events=None):
        """One-kernel production -> propagate -> decays.  With materialize=True also fills axion_energy/axion_flux/
        decay_axion_weight/scatter_axion_weight (device) and returns (rate tensor, event-weight tensor).
        hist_edges (array or CUDA tensor): also accumulate the event-weighted E_a spectrum (np.histogram semantics)
        in the same pass; it is added into `hist` (CUDA tensor) or returned as `self.last_hist`."""
        if self._rows is None:
            raise RuntimeError("simulate() must be called first")
        width, rescale = self._width_rescale(new_coupling)
        prop = self._prop_params(width, rescale, self._geom_accept() if self.is_isotropic else None, None)
        ev = P.event_params(days_exposure, threshold)
        rt = self._rt
        n = self._rows[1] * self._rows[2]
        with rt.activate():
            rate = rt.empty(1)
            hk = {}
            if hist_edges is not None:
                ed = hist_edges if torch.is_tensor(hist_edges) else rt.upload(np.asarray(hist_edges, dtype=np.float64))
                self.last_hist = hist if hist is not None else rt.zeros(int(ed.shape[0]) - 1)
                hk = dict(hist_edges=ed, hist=self.last_hist)
            if materialize:
                E, F = rt.empty(n), rt.empty(n)
                d32, s32 = rt.empty(n, torch.float32), rt.empty(n, torch.float32)
                evw = rt.empty(n)
                if events is not None:                      # (start, stop) CUDA events recorded right around the launch
                    events[0].record()
                self._launch_samples(E, F, prop, ev, d32=d32, s32=s32, evw=evw, rate=rate, **hk)
                if events is not None:
                    events[1].record()
                self._pending_production = None
                self._pending_prop = None
                self._set_production(E, F)
                self._D32, self._S32 = d32, s32
                return rate, evw
            if events is not None:
                events[0].record()
            self._launch_samples(None, None, prop, ev, rate=rate, **hk)
            if events is not None:
                events[1].record()
            return rate, None

    def _width_rescale(self, new_coupling):
        if new_coupling is not None:
            return self.decay_width(new_coupling, self.ma), power(new_coupling/self.ge, 2)
        return self.decay_width(self.ge, self.ma), 1.0

    def propagate(self, new_coupling=None):
        width, rescale = self._width_rescale(new_coupling)
        self._propagate(width, rescale, self._geom_accept() if self.is_isotropic else None)


class FluxComptonIsotropic(_ChainFlux):
    """
    ALP flux from Compton-like scattering gamma e- -> a e- (reference: fluxes.py:188-252).  `n_samples` uniform
    E_a draws per photon row, weighted by dsigma/dE_a / sigma_abs.
    """
    def __init__(self, photon_flux=[[1, 1]], target=None, det_dist=4.,
                 det_length=0.2, det_area=0.04, axion_mass=0.1, axion_coupling=1e-3, n_samples=100,
                 is_isotropic=True, loop_decay=False, **kwargs):
        target = Material("W") if target is None else target
        super().__init__(axion_mass, target, det_dist, det_length, det_area, **kwargs)
        self.photon_flux = photon_flux
        self.ge = axion_coupling
        self.n_samples = n_samples
        self.target_photon_xs = AbsCrossSection(target)
        self.is_isotropic = is_isotropic
        self.loop_decay = loop_decay
        self.row_offset = 0      # global index of photon_flux row 0 when the table is one shard of a larger one

    _SAMPLES_FN = "alpk_compton_samples"
    _TABLE_ATTR = "photon_flux"

    def _row_keep(self, table):
        """Rows that produce samples: fluxes.py:214-216 returns for s < (m_e + m_a)^2."""
        return ~(2 * M_E * table[:, 0] + M_E ** 2 < (M_E + self.ma)**2)

    def _stage_inputs(self, uniforms=None):
        """Host-side row filter (fluxes.py:214-216) + upload of the surviving rows, their global row ids (Philox
        key) and, in parity mode, the injected uniforms."""
        pf = _as_flux_table(self.photon_flux, "photon_flux")
        ns = int(self.n_samples)
        seed = self.seed if self.seed is not None else (_draw_seed() if uniforms is None else 0)
        self._seed_used = seed
        rt = self._rt
        with rt.activate():
            if pf.flags.c_contiguous:
                d_e, d_w, d_ids, n_rows = rt.stage_flux_rows(pf, 1, float((M_E + self.ma)**2), int(self.row_offset))
            else:
                s = 2 * M_E * pf[:, 0] + M_E ** 2
                keep = ~(s < (M_E + self.ma)**2)
                idx, e_k, w_k = _kept_columns(pf, keep)
                n_rows = int(idx.shape[0])
                d_e, d_w, d_ids = rt.upload_packed([(e_k, np.float64), (w_k, np.float64),
                                                    (idx + int(self.row_offset), np.uint32)])
            if uniforms is not None:
                uniforms = np.ascontiguousarray(uniforms, dtype=np.float64).ravel()
                if uniforms.shape[0] != n_rows * ns:
                    raise ValueError(f"uniforms must hold n_rows*n_samples = {n_rows * ns} variates, got {uniforms.shape[0]}")
            self._inputs = (d_e, d_w, n_rows, ns, d_ids, rt.upload(uniforms) if uniforms is not None else None, seed)
            self._table = rt.empty_block(n_rows * N.COMPTON_NP)
        return self._inputs

    def _launch_rows(self):
        """Per-row constants (threshold window, prefactors, sigma_abs lookup) -> row table, on the device."""
        d_eg, d_ph, n_rows, ns, d_ids, d_u, seed = self._inputs
        par = P.compton_params(self.ma, self.ge, self.target_z, ns, fast=self.precision != "faithful", rng_rounds=self.rng_rounds)
        rt = self._rt
        with rt.activate():
            if n_rows:
                le, lx, nt = self.target_photon_xs.device_table(rt)
                N.call("alpk_compton_rows", N.ptr(d_eg), N.ptr(d_ph), n_rows, N.ptr(le), N.ptr(lx), nt,
                       C.byref(par), N.ptr(self._table), rt.stream())
        self._rows = (self._table, n_rows, ns, d_ids, par, d_u, seed)

    def simulate(self, uniforms=None):
        """uniforms: optional U[0,1) array, n_samples per row passing the threshold, row-major (parity mode)."""
        self._reset_samples()
        self._stage_inputs(uniforms)
        self._launch_rows()
        self._finish_simulate()


_DEVICE_KW = ("seed", "device", "fused", "keep_f64", "precision", "strict")


def _split_kwargs(kwargs):
    """The reference's e+- flux classes swallow arbitrary **kwargs; keep only the device options."""
    return {k: kwargs[k] for k in _DEVICE_KW if k in kwargs}


def track_length_prob(Ei, Ef, t):
    """Track-length pdf (fluxes.py:257-259); host helper for scripts, the kernels carry their own copy."""
    from scipy.special import gamma
    b = 4/3
    return np.heaviside(Ei-Ef, 1.0) * abs(power(np.log(Ei/Ef), b*t - 1) / (Ei * gamma(b*t)))


class FluxBremIsotropic(_ChainFlux):
    """
    ALP bremsstrahlung flux from an electron (+ positron) shower flux (reference: fluxes.py:271-375).  With
    `simulate(use_track_length=True)` every electron row draws one track-length-attenuated energy E1 in
    (max(ma, me), E0) (weight: (E0-ep_min) (Phi_e+Phi_p)(E0) x the integrated track-length probability), then
    `n_samples` log-uniform E_a draws weighted by the Tsai (pseudoscalar) or IWW (vector) spectrum.
    """
    _SAMPLES_FN = "alpk_brem_samples"

    def __init__(self, electron_flux=[[1., 0.]], positron_flux=[1., 0.], target=None,
                 target_length=10.0, det_dist=4., det_length=0.2, det_area=0.04,
                 axion_mass=0.1, axion_coupling=1e-3, n_samples=100,
                 is_isotropic=True, loop_decay=False, boson_type="pseudoscalar",
                 max_track_length=5.0, is_monoenergetic=False, **kwargs):
        target = Material("W") if target is None else target
        super().__init__(axion_mass, target, det_dist, det_length, det_area, **_split_kwargs(kwargs))
        self.electron_flux = electron_flux
        self.positron_flux = positron_flux
        self.boson_type = boson_type
        self.ge = axion_coupling
        self.target_density = target.density  # g/cm3
        self.target_radius = target_length  # cm
        self.ntargets_by_area = target_length * target.density * AVOGADRO / (2*target.z[0])  # N_T / cm^2
        self.ntarget_area_density = target.rad_length * AVOGADRO / (2*target.z[0])
        self._rad_length = target.rad_length
        self.n_samples = n_samples
        self.is_isotropic = is_isotropic
        self.loop_decay = loop_decay
        self.max_t = max_track_length
        self.is_monoenergetic = is_monoenergetic
        self.row_offset = 0

    def electron_flux_dN_dE(self, energy):
        ef = np.asarray(self.electron_flux, dtype=np.float64)
        return np.interp(energy, ef[:, 0], ef[:, 1], left=0.0, right=0.0)

    def positron_flux_dN_dE(self, energy):
        pf = np.asarray(self.positron_flux, dtype=np.float64)
        return np.interp(energy, pf[:, 0], pf[:, 1], left=0.0, right=0.0)

    def simulate(self, use_track_length=True, uniforms=None):
        """uniforms (parity mode): flat U[0,1) array in the reference's draw order -- per electron row with
        E0 >= max(ma, me) one scalar, followed by `n_samples` variates iff that row emits (fluxes.py:348-352,322-326);
        without track length: `n_samples` per emitting row."""
        self._reset_samples()
        ef = _as_flux_table(self.electron_flux, "electron_flux")
        ns = int(self.n_samples)
        par = P.brem_params(self.ma, self.ge, self.target_z, self._rad_length, ns, self.boson_type, self.max_t,
                            use_track_length, fast=self.precision != "faithful", rng_rounds=self.rng_rounds)
        if use_track_length:
            ep_min = max(self.ma, M_E)                                           # fluxes.py:346
            keep = ~(ef[:, 0] < ep_min)
            if self.is_monoenergetic:
                if ef.shape[0] != 1:
                    # the reference multiplies an (n_rows,) weight array into (n_samples,) samples (fluxes.py:310-314,337)
                    raise ValueError("operands could not be broadcast together: is_monoenergetic needs a single-row electron_flux")
                ele = ef[:, :2].copy()
                pos = np.array([[ef[0, 0], 0.0]])
            else:
                ele = ef[:, :2]
                pos = np.asarray(self.positron_flux, dtype=np.float64)
                if pos.ndim != 2:
                    # reference: positron_flux[:,0] on the default 1-D [1., 0.] (fluxes.py:306-307)
                    raise TypeError("list indices must be integers or slices, not tuple: positron_flux must have shape (n, 2)")
        else:
            keep = np.ones(ef.shape[0], dtype=bool)
            ele = pos = np.zeros((1, 2))
        rows = np.nonzero(keep)[0]
        e0, w0 = ef[keep, 0], ef[keep, 1]
        n_rows = int(rows.shape[0])
        seed = self.seed if self.seed is not None else (_draw_seed() if uniforms is None else 0)
        self._seed_used = seed

        u_rows = u_samp = None
        if uniforms is not None:
            u_rows, u_samp = self._split_uniforms(np.asarray(uniforms, dtype=np.float64).ravel(), e0, ns, use_track_length)

        rt = self._rt
        gx, gw = P.gauss_legendre()
        with rt.activate():
            table = rt.empty_block(n_rows * N.BREM_NP)
            emit = torch.zeros(n_rows, dtype=torch.uint8, device=rt.device)
            ids = (rows + int(self.row_offset)).astype(np.uint32)
            # ONE pinned H2D copy for the row inputs, both shower tables, the Gauss-Legendre nodes and the row ids
            packed = rt.upload_packed([(a, np.float64) for a in (e0, w0, ele[:, 0], ele[:, 1], pos[:, 0], pos[:, 1], gx, gw)]
                                      + [(ids, np.uint32)])
            d, d_ids = packed[:8], packed[8]
            if n_rows:
                d_ur = rt.upload(u_rows) if u_rows is not None else None
                N.call("alpk_brem_rows", N.ptr(d[0]), N.ptr(d[1]), n_rows, N.ptr(d_ids), N.ptr(d_ur), C.c_uint64(seed),
                       N.ptr(d[2]), N.ptr(d[3]), int(ele.shape[0]), N.ptr(d[4]), N.ptr(d[5]), int(pos.shape[0]),
                       N.ptr(d[6]), N.ptr(d[7]), C.byref(par), N.ptr(table), N.ptr(emit), rt.stream())
            # rows whose drawn energy cannot radiate (ea_max <= ma) emit nothing (fluxes.py:322-324): compact the table
            sel = torch.nonzero(emit, as_tuple=False).ravel()
            n_emit = int(sel.shape[0])                                           # (device -> host: one small sync)
            ctable = table.view(n_rows, N.BREM_NP).index_select(0, sel).contiguous().view(-1) if n_rows else table
            cids = d_ids.index_select(0, sel).contiguous() if n_rows else d_ids
            d_u = rt.upload(u_samp) if u_samp is not None else None
        if u_samp is not None and u_samp.shape[0] != n_emit * ns:
            raise ValueError("injected uniforms do not match the rows that emit")
        self._row_table_all = table
        self._emit = emit
        self._rows = (ctable, n_emit, ns, cids, par, d_u, seed)
        self._finish_simulate()

    def _split_uniforms(self, u, e0, ns, use_track_length):
        """Split the reference-ordered flat stream into the per-row scalars and the per-sample block."""
        if not use_track_length:
            return None, np.ascontiguousarray(u)
        ep_min = max(self.ma, M_E)
        u_rows = np.zeros(e0.shape[0])
        samp = []
        pos = 0
        for k in range(e0.shape[0]):
            if pos >= u.shape[0]:
                raise ValueError("uniforms exhausted")
            u_rows[k] = u[pos]; pos += 1
            e1 = ep_min + (e0[k] - ep_min) * u_rows[k]                           # fluxes.py:350, same IEEE ops as the kernel
            ea_max = e1 * (1 - (self.ma / e1) * (self.ma / e1))
            if ea_max > self.ma:
                samp.append(u[pos:pos + ns]); pos += ns
        if pos != u.shape[0]:
            raise ValueError(f"uniforms hold {u.shape[0]} variates, the reference draw order consumes {pos}")
        return u_rows, (np.ascontiguousarray(np.concatenate(samp)) if samp else np.zeros(0))


class FluxResonanceIsotropic(AxionFlux):
    """
    Resonant e+ e- -> a production from a positron flux (reference: fluxes.py:380-464): a single ALP sample at
    E_a = ma^2/(2 me) whose weight is an `n_samples` Monte-Carlo integral over (E+, t) of flux x track-length pdf.
    """
    def __init__(self, positron_flux=[[1., 0.]], target=None, target_length=10.0,
                 det_dist=4., det_length=0.2, det_area=0.04, axion_mass=0.1, axion_coupling=1e-3,
                 n_samples=100, is_isotropic=True, loop_decay=False, boson_type="pseudoscalar",
                 max_track_length=5.0, **kwargs):
        target = Material("W") if target is None else target
        super().__init__(axion_mass, target, det_dist, det_length, det_area, n_samples, **_split_kwargs(kwargs))
        self.positron_flux = positron_flux
        pf = np.asarray(positron_flux, dtype=np.float64)
        self.positron_flux_bin_widths = pf[1:, 0] - pf[:-1, 0]
        self.ge = axion_coupling
        self.target_radius = target_length  # cm
        self.ntarget_area_density = target.rad_length * AVOGADRO / (target.z[0] + target.n[0])  # fluxes.py:397
        self.is_isotropic = is_isotropic
        self.loop_decay = loop_decay
        self.boson_type = boson_type
        self.max_t = max_track_length

    decay_width = _width_e
    propagate_iso_vol_int = _vol_int_e

    def positron_flux_dN_dE(self, energy):
        pf = np.asarray(self.positron_flux, dtype=np.float64)
        return np.interp(energy, pf[:, 0], pf[:, 1], left=0.0, right=0.0)

    def resonance_peak(self):
        if self.boson_type == "pseudoscalar":
            return 2*pi*M_E*power(self.ge / self.ma, 2) / sqrt(1 - power(2*M_E/self.ma, 2))
        elif self.boson_type == "vector":
            return 12*pi*self.ge**2 * ALPHA / M_E / 6

    def simulate(self, uniforms=None):
        """uniforms (parity mode): [2, n_samples] -- the E+ draws then the t draws (fluxes.py:437-438)."""
        self._reset_samples()
        pf = _as_flux_table(self.positron_flux, "positron_flux")
        resonant_energy = -M_E + self.ma**2 / (2 * M_E)                          # fluxes.py:427
        # the reference takes Python's max() of the column (fluxes.py:426): 0.4 ms of interpreter time for 1e4 rows -- the vectorised
        # maximum is the same number unless the column holds NaN (then keep the builtin's order-dependent answer)
        emax = pf[:, 0].max()
        if emax != emax:
            emax = max(pf[:, 0])
        if resonant_energy + M_E < self.ma or resonant_energy < M_E or resonant_energy > emax:   # :428-435
            return
        ns = int(self.n_samples)
        seed = self.seed if self.seed is not None else (_draw_seed() if uniforms is None else 0)
        self._seed_used = seed
        rt = self._rt
        with rt.activate():
            d_u = None
            if uniforms is not None:
                u = np.ascontiguousarray(uniforms, dtype=np.float64).ravel()
                if u.shape[0] != 2 * ns:
                    raise ValueError(f"uniforms must hold 2*n_samples = {2 * ns} variates")
                d_u = rt.upload(u)
            d_e, d_f = rt.upload_packed([(pf[:, 0], np.float64), (pf[:, 1], np.float64)])
            total = rt.empty(1)
            N.call("alpk_resonance_sum", N.ptr(d_e), N.ptr(d_f), int(pf.shape[0]), float(resonant_energy), float(emax),
                   float(self.max_t), ns, N.ptr(d_u), C.c_uint64(seed), 0 if self.precision == "faithful" else 1,
                   N.ptr(total), N.ptr(rt.scratch), rt.stream())
            s = float(total.item())
        mc_vol = self.max_t*(emax - resonant_energy)                             # :439
        attenuated_flux = mc_vol*s/self.n_samples                                # :441
        wgt = self.target_z * (self.ntarget_area_density * HBARC**2) * self.resonance_peak() * attenuated_flux   # :442
        self.axion_energy = [self.ma**2 / (2 * M_E)]                             # :444
        self.axion_flux = [wgt]                                                  # :445

    def propagate(self, new_coupling=None):
        if new_coupling is not None:
            width, rescale = self.decay_width(new_coupling, self.ma), power(new_coupling/self.ge, 2)
        else:
            width, rescale = self.decay_width(self.ge, self.ma), 1.0
        self._propagate(width, rescale, self._geom_accept() if self.is_isotropic else None)


class FluxPairAnnihilationIsotropic(_ChainFlux):
    """
    Associated production e+ e- -> a gamma from a positron flux on atomic electrons at rest (reference:
    fluxes.py:469-549): per positron row above threshold, `n_samples` isotropic CM-frame draws weighted by the
    tree-level matrix element (1% soft-photon cut), boosted to lab-frame ALP energies.
    """
    _SAMPLES_FN = "alpk_pairann_samples"

    def __init__(self, positron_flux=[[1., 0.]], target=None,
                 det_dist=4., det_length=0.2, det_area=0.04, axion_mass=0.1,
                 axion_coupling=1e-3, n_samples=100, is_isotropic=True,
                 loop_decay=False, **kwargs):
        target = Material("W") if target is None else target
        super().__init__(axion_mass, target, det_dist, det_length, det_area, n_samples, **_split_kwargs(kwargs))
        self.positron_flux = positron_flux
        pf = np.asarray(positron_flux, dtype=np.float64)
        self.positron_flux_bin_widths = pf[1:, 0] - pf[:-1, 0]
        self.ge = axion_coupling
        self.ntarget_area_density = target.rad_length * AVOGADRO / (2*target.z[0])
        self._rad_length = target.rad_length
        self.is_isotropic = is_isotropic
        self.loop_decay = loop_decay
        self.row_offset = 0

    def p1_cm(self, s):
        return np.sqrt((np.power(s - 2*M_E**2, 2) - np.power(2*M_E**2, 2))/(4*s))

    def p3_cm(self, s):
        return np.sqrt((np.power(s - self.ma**2, 2))/(4*s))

    def simulate(self, uniforms=None):
        """uniforms (parity mode): [n_rows, 2, n_samples] for the rows above threshold -- per row the phi draws then
        the theta draws (cross_section_mc.py:204-205)."""
        self._reset_samples()
        pf = _as_flux_table(self.positron_flux, "positron_flux")
        keep = ~(pf[:, 0] < max((self.ma**2 - M_E**2)/(2*M_E), M_E))            # fluxes.py:507
        rows = np.nonzero(keep)[0]
        n_rows, ns = int(rows.shape[0]), int(self.n_samples)
        seed = self.seed if self.seed is not None else (_draw_seed() if uniforms is None else 0)
        self._seed_used = seed
        par = P.pairann_params(self.ma, self.ge, self.target_z, self._rad_length, ns, fast=self.precision != "faithful",
                               rng_rounds=self.rng_rounds)
        rt = self._rt
        with rt.activate():
            d_u = None
            if uniforms is not None:
                u = np.ascontiguousarray(uniforms, dtype=np.float64).ravel()
                if u.shape[0] != 2 * n_rows * ns:
                    raise ValueError(f"uniforms must hold 2*n_rows*n_samples = {2 * n_rows * ns} variates")
                d_u = rt.upload(u)
            table = rt.empty_block(n_rows * N.PAIRANN_NP)
            d_e, d_w, d_ids = rt.upload_packed([(pf[keep, 0], np.float64), (pf[keep, 1], np.float64),
                                                ((rows + int(self.row_offset)).astype(np.uint32), np.uint32)])   # one pinned H2D copy
            if n_rows:
                N.call("alpk_pairann_rows", N.ptr(d_e), N.ptr(d_w), n_rows, C.byref(par), N.ptr(table), rt.stream())
        self._rows = (table, n_rows, ns, d_ids, par, d_u, seed)
        self._finish_simulate()


class FluxNuclearIsotropic(AxionFlux):
    """
    ALP flux from nuclear de-excitation lines: rows of [E_transition, decays/s, beta, eta] times the a/gamma branching
    ratio (reference: fluxes.py:554-600).
    """
    def __init__(self, transition_rates=np.array([[1.0, 0.0, 1.0, 0.5]]), target=None,
                 det_dist=4., det_length=0.2, det_area=0.04, is_isotropic=True,
                 axion_mass=0.1, gann0=1e-3, gann1=1e-3, gae=0.0, gagamma=0.0, n_samples=100, **kwargs):
        target = Material("W") if target is None else target
        super().__init__(axion_mass, target, det_dist, det_length, det_area, n_samples, **_split_kwargs(kwargs))
        self.rates = transition_rates
        self.gann0 = gann0
        self.gann1 = gann1
        self.gae = gae
        self.gagamma = gagamma
        self.is_isotropic = is_isotropic

    def decay_width(self):
        return W_ee(self.gae, self.ma) + W_ee(self.gagamma, self.ma)

    def simulate(self, j=1, delta=0.0):
        self._reset_samples()
        rates = np.asarray(self.rates, dtype=np.float64)
        keep = rates[:, 0] > self.ma                                              # fluxes.py:588
        rows = np.ascontiguousarray(rates[keep, :4])
        n = int(rows.shape[0])
        c0 = (j/(j+1)) / (1 + delta**2) / pi / ALPHA                              # :574
        rt = self._rt
        with rt.activate():
            E, F = rt.empty(n), rt.empty(n)
            if n:
                d_rows = rt.upload(rows.ravel())
                N.call("alpk_nuclear", N.ptr(d_rows), n, float(self.ma**2), float(c0), float(2*j + 1),
                       float(self.gann0), float(self.gann1), N.ptr(E), N.ptr(F), rt.stream())
        self._set_production(E, F)

    def propagate(self):
        self._propagate(W_ee(self.gae, self.ma), 1.0, self._geom_accept() if self.is_isotropic else None)

    def propagate_iso_vol_int(self, geom: DetectorGeometry):
        AxionFlux.propagate_iso_vol_int(self, geom, self.decay_width())


class FluxPi0Isotropic(AxionFlux):
    """
    ALPs (or dark photons) from pi0 -> gamma + a (reference: fluxes.py:926-991): `simulate()` for pi0 at rest,
    `simulate_flux(momenta)` for a list of pi0 momenta along z with optional energy / angle cuts.  `propagate()` keeps
    float64 weights (no decays, acceptance only), as the reference does for this class.
    """
    def __init__(self, pi0_rate=0.0259, boson_mass=0.1, coupling=1.0,
                 det_dist=20.0, det_area=2.0, det_length=1.0,
                 target=None, n_samples=1000, **kwargs):
        from .constants import M_PI0
        target = Material("W") if target is None else target
        super().__init__(boson_mass, target, det_dist, det_length, det_area, **_split_kwargs(kwargs))
        self.meson_rate = pi0_rate
        self.n_samples = n_samples
        self.g = coupling
        self.boson_mass = boson_mass
        self.mm = M_PI0

    def br(self):
        return 2 * (self.g)**2 * (1 - power(self.ma / self.mm, 2))**3 / sqrt(4*pi*ALPHA)

    def _cm(self):
        p_cm = (self.mm**2 - self.ma**2)/(2*self.mm)
        return p_cm, sqrt(p_cm**2 + self.ma**2)

    def simulate_flux(self, pi0_flux, energy_cut=0.0, angle_cut=np.pi, uniforms=None):
        """pi0_flux: array of pi0 momenta [MeV], normalised to the pi0 rate.  uniforms: U[0,1) per momentum (parity)."""
        self._reset_samples()
        if self.ma > self.mm:
            return
        p = np.ascontiguousarray(pi0_flux, dtype=np.float64).ravel()
        n = int(p.shape[0])
        p_cm, e1_cm = self._cm()
        seed = self.seed if self.seed is not None else (_draw_seed() if uniforms is None else 0)
        rt = self._rt
        with rt.activate():
            E, F = rt.empty(n), rt.empty(n)
            keep = torch.zeros(n, dtype=torch.uint8, device=rt.device)
            if n:
                d_u = rt.upload(np.asarray(uniforms, dtype=np.float64).ravel()) if uniforms is not None else None
                d_p = rt.upload(p)
                N.call("alpk_pi0_flux", N.ptr(d_p), n, N.ptr(d_u), C.c_uint64(seed), float(self.mm**2),
                       float(self.ma**2), float(p_cm), float(e1_cm), float(energy_cut), float(angle_cut),
                       float(self.meson_rate*self.br()/n), N.ptr(E), N.ptr(F), N.ptr(keep), rt.stream())
                sel = torch.nonzero(keep, as_tuple=False).ravel()
                E, F = E.index_select(0, sel).contiguous(), F.index_select(0, sel).contiguous()
        self._set_production(E, F)

    def simulate(self):
        self._reset_samples()
        if self.ma > self.mm:
            return
        p_cm, e1_cm = self._cm()
        ns = int(self.n_samples)
        rt = self._rt
        with rt.activate():                       # constant arrays (fluxes.py:982-983): device fills, no arithmetic
            E = torch.full((ns,), float(e1_cm), dtype=torch.float64, device=rt.device)
            F = torch.full((ns,), float(self.meson_rate * self.br() * 1.0 / np.sum(self.n_samples)), dtype=torch.float64,
                           device=rt.device)
        self._set_production(E, F)

    def propagate(self, is_isotropic=True):
        self._ensure_production()
        wgt = self._F if self._F is not None else torch.zeros(0, dtype=torch.float64, device=self._rt.device)
        geom = self.det_area / (4*pi*self.det_dist**2)
        rt = self._rt
        n = int(wgt.shape[0])
        with rt.activate():                       # float64 weights for this class (fluxes.py:985-991)
            d, s_ = rt.empty(n), rt.empty(n)
            N.call("alpk_scale", N.ptr(wgt), n, 0.0, N.ptr(d), rt.stream())
            N.call("alpk_scale", N.ptr(wgt), n, float(geom) if is_isotropic else 1.0, N.ptr(s_), rt.stream())
        self._D32, self._S32 = d, s_
        for k in ("D32", "S32"):
            self._host.pop(k, None)


class FluxPairAnnihilationGamma(AxionFlux):
    """
    Associated production e+ e- -> gamma a through a virtual photon (photon coupling) from a positron flux, with
    track-length smearing of the positron energy (reference: fluxes.py:1058-1148).  Per positron-table bin centre above
    threshold: `n_samples` (depth, smeared energy) draws, then one E_a draw per smeared positron.
    """
    def __init__(self, positron_flux=[1., 0.], target=None,
                 target_radiation_length=6.76, det_dist=4., det_length=0.2, det_area=0.04,
                 axion_mass=0.1, axion_coupling=1e-3, n_samples=100, is_isotropic=True, **kwargs):
        target = Material("W") if target is None else target
        super().__init__(axion_mass, target, det_dist, det_length, det_area, n_samples, **_split_kwargs(kwargs))
        self.positron_flux = positron_flux
        pf = np.asarray(positron_flux, dtype=np.float64)
        self.positron_flux_bin_widths = pf[1:, 0] - pf[:-1, 0]
        self.positron_flux_bin_centers = (pf[1:, 0] + pf[:-1, 0])/2
        self.gagamma = axion_coupling
        self.ntarget_area_density = target_radiation_length * AVOGADRO / (2*target.z[0])
        self.is_isotropic = is_isotropic
        self.row_offset = 0

    def positron_flux_dN_dE(self, energy):
        pf = np.asarray(self.positron_flux, dtype=np.float64)
        return np.interp(energy, pf[:, 0], pf[:, 1], left=0.0, right=0.0)

    def simulate(self, uniforms=None):
        """uniforms (parity mode): flat stream in the reference's draw order -- per bin above threshold `n_samples`
        depths then `n_samples` smeared energies (fluxes.py:1099-1100), followed by one E_a draw per sample (:1126)."""
        self._reset_samples()
        pf = _as_flux_table(self.positron_flux, "positron_flux")
        ma, ns = self.ma, int(self.n_samples)
        ep_min = max((ma**2 - M_E**2)/(2*M_E), M_E)                              # fluxes.py:1093
        keep = ~(self.positron_flux_bin_centers < ep_min)                        # :1096
        bins = np.nonzero(keep)[0]
        nb = int(bins.shape[0])
        par = N.PairGammaT()
        par.ma, par.ma2, par.ma4 = float(ma), float(ma**2), float(ma**4)
        par.me2, par.me4 = float(M_E**2), float(M_E**4)
        par.ep_min = float(ep_min)
        par.nt = float(self.ntarget_area_density * HBARC**2)
        par.zc = float(self.target_z * 4*pi*ALPHA*self.gagamma**2)
        par.n_samples, par.max_t = float(ns), 5.0
        par.fast = 0 if self.precision == "faithful" else 1
        seed = self.seed if self.seed is not None else (_draw_seed() if uniforms is None else 0)
        rt = self._rt
        n = nb * ns
        with rt.activate():
            E, F = rt.empty(n), rt.empty(n)
            if n:
                d_ub = d_ua = None
                if uniforms is not None:
                    u = np.ascontiguousarray(uniforms, dtype=np.float64).ravel()
                    if u.shape[0] != 3 * n:
                        raise ValueError(f"uniforms must hold 3*n_bins*n_samples = {3 * n} variates")
                    d_ub, d_ua = rt.upload(u[:2 * n]), rt.upload(u[2 * n:])
                # (keep every uploaded tensor referenced until the launch: temporaries would be recycled by the allocator)
                d_c, d_w = rt.upload(self.positron_flux_bin_centers[keep]), rt.upload(self.positron_flux_bin_widths[keep])
                d_ids = rt.upload((bins + int(self.row_offset)).astype(np.uint32), np.uint32)
                d_pe, d_pf = rt.upload(pf[:, 0]), rt.upload(pf[:, 1])
                N.call("alpk_pairgamma", N.ptr(d_c), N.ptr(d_w), nb, ns, N.ptr(d_ids), N.ptr(d_pe), N.ptr(d_pf),
                       int(pf.shape[0]), C.byref(par), N.ptr(d_ub), N.ptr(d_ua), C.c_uint64(seed), N.ptr(E), N.ptr(F),
                       rt.stream())
        self._set_production(E, F)

    def propagate(self, new_coupling=None):
        geom = self._geom_accept() if self.is_isotropic else None
        if new_coupling is not None:
            self._propagate(W_gg(new_coupling, self.ma), power(new_coupling/self.gagamma, 2), geom)
        else:
            self._propagate(W_gg(self.gagamma, self.ma), 1.0, geom)


class FluxNeutralMeson2BodyDecay(AxionFlux):
    """
    ALP / dark-photon flux from neutral mesons decaying in flight, meson -> gamma + a (reference: fluxes.py:996-1052).
    `meson_flux`: rows (px, py, pz, E).  Every meson is decayed `n_samples` times with the general two-body decay kernel
    (dense Lorentz boost); ALP energies, polar angles and weights are appended to `axion_energy` / `axion_angle` /
    `axion_flux` -- like the reference, `simulate()` does not clear earlier results.  Weights stay float64.
    """
    def __init__(self, meson_flux, flux_weight, boson_mass=0.1, coupling=1.0, meson_mass=None,
                 det_dist=20.0, det_area=2.0, det_length=1.0, apply_angle_cut=False,
                 n_samples=10, off_axis_angle=0.0, **kwargs):
        from .constants import M_PI0
        super().__init__(boson_mass, Material("Be"), det_dist, det_length, det_area, **_split_kwargs(kwargs))
        self.meson_flux = meson_flux
        self.m_meson = M_PI0 if meson_mass is None else meson_mass
        self.n_samples = n_samples
        self.g = coupling
        self.boson_mass = boson_mass
        self.mm = M_PI0
        self.apply_angle_cut = apply_angle_cut
        self.flux_weight = flux_weight
        self.off_axis_angle = off_axis_angle
        self.phi_range = 2*pi
        if off_axis_angle > 0.0:
            self.phi_range = 2*self.det_sa()
        self._A = None                    # device float64: axion_angle

    @property
    def axion_angle(self):
        a = getattr(self, "_A", None)
        return [] if a is None else self._host_get("A", a, [])

    @axion_angle.setter
    def axion_angle(self, v):
        if isinstance(v, list) and len(v) == 0:
            self._A = None
        else:
            self._A = self._rt.upload(np.asarray(v, dtype=np.float64).ravel())
        if hasattr(self, "_host"):
            self._host.pop("A", None)

    def br(self):
        if self.ma > self.m_meson:
            return 0.0
        return 2 * (self.g)**2 * (1 - power(self.ma / self.m_meson, 2))**3 / (4*pi*ALPHA)

    def simulate(self, uniforms=None):
        """uniforms (parity mode): [n_mesons, 2, n_samples] -- per meson the phi draws then the theta draws."""
        mes = np.ascontiguousarray(self.meson_flux, dtype=np.float64)
        if mes.ndim != 2 or mes.shape[1] < 4:
            raise IndexError("meson_flux must have shape (n, 4): rows of [px, py, pz, E]")
        nm, ns = int(mes.shape[0]), int(self.n_samples)
        n = nm * ns
        w = self.br() * (1.0 / ns) * self.flux_weight                            # fluxes.py:1031 (mc.weights = 1/n)
        if self.apply_angle_cut:
            w_final = w * self.phi_range/(2*pi)                                   # :1037
        else:
            w_final = self.det_area * w / (4*pi*self.det_dist**2)                 # :1040
        seed = self.seed if self.seed is not None else (_draw_seed() if uniforms is None else 0)
        rt = self._rt
        with rt.activate():
            p4 = rt.upload(np.ascontiguousarray(mes[:, [3, 0, 1, 2]]).ravel())    # (E, px, py, pz)  :1027
            table = rt.empty(nm * N.PARENT_GENERAL_NP)
            p1 = torch.empty((4, n), dtype=torch.float64, device=rt.device)
            theta = rt.empty(n)
            d_u = None
            if uniforms is not None:
                u = np.ascontiguousarray(uniforms, dtype=np.float64).ravel()
                if u.shape[0] != 2 * n:
                    raise ValueError(f"uniforms must hold 2*n_mesons*n_samples = {2 * n} variates")
                d_u = rt.upload(u)
            if n:
                N.call("alpk_decay2body_general", N.ptr(p4), None, nm, float(self.boson_mass), 0.0, ns, None, N.ptr(d_u),
                       C.c_uint64(seed), N.ptr(table), N.ptr(p1), None, None, N.ptr(theta), rt.stream())
            E = p1[0].contiguous()
            if self.apply_angle_cut:
                sa = float(self.det_sa())
                keep = (theta < sa + self.off_axis_angle) & (theta > self.off_axis_angle - sa)        # :1034-1035
                sel = torch.nonzero(keep, as_tuple=False).ravel()
                E, theta = E.index_select(0, sel), theta.index_select(0, sel)
            F = torch.full((int(E.shape[0]),), float(w_final), dtype=torch.float64, device=rt.device)
            if self._E is not None and int(self._E.shape[0]):                    # the reference keeps appending
                E, F, theta = torch.cat([self._E, E]), torch.cat([self._F, F]), torch.cat([self._A, theta])
        self._set_production(E.contiguous(), F.contiguous())
        self._A = theta.contiguous()
        self._host.pop("A", None)

    def propagate(self):
        self._ensure_production()
        rt = self._rt
        wgt = self._F if self._F is not None else rt.zeros(0)
        n = int(wgt.shape[0])
        with rt.activate():                       # float64 weights (fluxes.py:1048-1051)
            d, s_ = rt.empty(n), rt.empty(n)
            N.call("alpk_scale", N.ptr(wgt), n, 0.0, N.ptr(d), rt.stream())
            N.call("alpk_scale", N.ptr(wgt), n, 1.0, N.ptr(s_), rt.stream())
        self._D32, self._S32 = d, s_
        for k in ("D32", "S32"):
            self._host.pop(k, None)

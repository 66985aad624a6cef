"""
ALP event generators -- drop-in for the reference's `generators.py` classes on the hot path
(ElectronEventGenerator generators.py:18-52, PhotonEventGenerator generators.py:57-128).

`decays`, `compton`, `inverse_primakoff` return the total event count (np.float64, like `np.sum`); the per-sample
event weights stay on the GPU and are copied to the host when the `decay_weights` / `scatter_weights` attributes are
read.  `simulate_decay_4vectors` returns two `LorentzVectorArray`s and a weight array (sequence-compatible with the
reference's three lists).
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import params as P
from .constants import METER_BY_MEV
from .fluxes import AxionFlux, _draw_seed
from .materials import Material
from .photon_xs import ALPPairProdutionCrossSection
from .vectors import LorentzVectorArray


class _EventGenerator:
    def __init__(self, flux: AxionFlux, detector: Material):
        self.flux = flux
        self.det_z = detector.z[0]
        self.efficiency = None
        self.energy_threshold = None
        self._dev = {}            # name -> device tensor of per-sample event weights
        self._hostcache = {}
        self.materialize_weights = True   # keep per-sample event weights (device) so the attributes can be read
        self._last_decay = None           # (event params, flux epoch) of the last decays(): lazily computed decay_weights

    # reference attributes (host views, lazily copied)
    @property
    def axion_energy(self):
        return np.array(self.flux.axion_energy)

    def _weights(self, name, like):
        t = self._dev.get(name)
        if t is None:
            return np.zeros_like(like)
        h = self._hostcache.get(name)
        if h is None:
            h = self._hostcache[name] = self.flux._rt.to_host(t)
        return h

    @property
    def decay_weights(self):
        """days*S_PER_DAY*decay_axion_weight*heaviside(E - threshold, 1) of the last decays() call (generators.py:88-90).
        In the reduce-only schedule (flux.fused=True, or decays_device()) the kernel returns only the rate; the per-sample
        weights are then computed here, on demand, by one alpk_decays launch on the (re)materialised flux arrays."""
        if self._dev.get("decay") is None and self._last_decay is not None:
            ev, epoch = self._last_decay
            fl = self.flux
            if epoch != fl._epoch:
                raise RuntimeError("decay_weights: the flux was re-simulated / re-propagated after decays(); call decays() again")
            fl._ensure_production()
            fl._ensure_propagated()
            rt = fl._rt
            n = 0 if fl._E is None else int(fl._E.shape[0])
            with rt.activate():
                rate, evw = rt.empty(1), rt.empty(n)
                N.call("alpk_decays", N.ptr(fl._E), N.ptr(fl._D32), n, C.byref(ev), N.ptr(evw), N.ptr(rate),
                       N.ptr(rt.scratch), rt.stream())
            self._store("decay", evw)
        return self._weights("decay", self.flux.decay_axion_weight)

    @property
    def scatter_weights(self):
        return self._weights("scatter", self.flux.scatter_axion_weight)

    @property
    def pair_weights(self):
        return self._weights("pair", self.flux.scatter_axion_weight)

    @staticmethod
    def _need_f32(t, what):
        if t is not None and t.dtype != torch.float32:
            raise TypeError(f"{what} is float64: flux classes that keep float64 weights (FluxPi0Isotropic) feed the "
                            "dark-sector generators, which are outside this path")

    def _store(self, name, t):
        self._dev[name] = t
        self._hostcache.pop(name, None)

    def decays(self, days_exposure, threshold):
        """days*S_PER_DAY * decay_axion_weight * heaviside(E - threshold, 1), summed (generators.py:48-52, 88-92)."""
        ev = P.event_params(days_exposure, threshold)
        self._need_f32(self.flux._D32, "decay_axion_weight")       # (None while a fused chain is still pending)
        rate, evw = self.flux._decay_rate(ev, self.materialize_weights and self.flux.fused is not True)
        self._store("decay", evw)
        self._last_decay = (ev, self.flux._epoch)
        return np.float64(rate.item())

    def decays_device(self, days_exposure, threshold):
        """Same as decays() but returns the rate as a 1-element CUDA tensor without synchronising."""
        ev = P.event_params(days_exposure, threshold)
        rate, evw = self.flux._decay_rate(ev, False)
        self._store("decay", None)                 # (no stale weights of an earlier call; computed on demand if read)
        self._last_decay = (ev, self.flux._epoch)
        return rate

    def decays_async(self, days_exposure, threshold):
        """decays() without the wait: queues the kernels and the device -> host copy of the rate and returns a
        runtime.PendingValue; `.result()` (or float()) gives what decays() returns.  A loop over many jobs that reads each
        rate one job late keeps the GPU busy while the host prepares the next job."""
        from .runtime import PendingValue
        rate = self.decays_device(days_exposure, threshold)
        with self.flux._rt.activate():
            return PendingValue(rate)

    def _scatter(self, kind, xs, ntargets, days_exposure, threshold):
        fl = self.flux
        fl._ensure_production()
        fl._ensure_propagated()
        rt = fl._rt
        n = 0 if fl._E is None else int(fl._E.shape[0])
        if fl._S32 is None or int(fl._S32.shape[0]) != n:
            raise ValueError("propagate() must run first: scatter_axion_weight does not match axion_energy")
        self._need_f32(fl._S32, "scatter_axion_weight")
        ev = P.event_params(days_exposure, threshold)
        if fl.precision != "faithful":
            kind += 2                               # fast arithmetic mode of the cross section (include/alpk.h)
        with rt.activate():
            rate = rt.empty(1)
            evw = rt.empty(n) if self.materialize_weights else None
            N.call("alpk_scatter_events", kind, N.ptr(fl._E), N.ptr(fl._S32), n, xs, float(ntargets / fl.det_area),
                   float(METER_BY_MEV**2), C.byref(ev), N.ptr(evw), N.ptr(rate), N.ptr(rt.scratch), rt.stream())
        self._store("scatter", evw)
        return np.float64(rate.item())


class ElectronEventGenerator(_EventGenerator):
    """ALP-electron coupling: decay, inverse-Compton and pair-production rates (generators.py:18-52)."""

    def __init__(self, flux: AxionFlux, detector: Material):
        super().__init__(flux, detector)
        self._detector = detector
        self._pair_xs = None

    @property
    def pair_xs(self):
        """ALPPairProdutionCrossSection(detector, ma=flux.ma) (generators.py:31), built when first used."""
        if self._pair_xs is None:
            self._pair_xs = ALPPairProdutionCrossSection(self._detector, ma=self.flux.ma)
        return self._pair_xs

    @pair_xs.setter
    def pair_xs(self, v):
        self._pair_xs = v

    def pair_production(self, ge, ntargets, days_exposure, threshold):
        """days*S_PER_DAY*(ntargets/area)*ge^2*sigma_pair(E)*METER_BY_MEV^2*scatter_w*heaviside (generators.py:33-39)."""
        fl = self.flux
        fl._ensure_production()
        fl._ensure_propagated()
        rt = fl._rt
        n = 0 if fl._E is None else int(fl._E.shape[0])
        if fl._S32 is None or int(fl._S32.shape[0]) != n:
            raise ValueError("propagate() must run first: scatter_axion_weight does not match axion_energy")
        ev = P.event_params(days_exposure, threshold)
        lead = days_exposure * P.S_PER_DAY * (ntargets / fl.det_area) * np.power(ge, 2)
        with rt.activate():
            le, lx, nt = self.pair_xs.device_table(rt)
            rate = rt.empty(1)
            evw = rt.empty(n) if self.materialize_weights else None
            N.call("alpk_pair_events", N.ptr(fl._E), N.ptr(fl._S32), n, N.ptr(le), N.ptr(lx), nt, float(lead),
                   float(METER_BY_MEV**2), C.byref(ev), N.ptr(evw), N.ptr(rate), N.ptr(rt.scratch), rt.stream())
        self._store("pair", evw)
        return np.float64(rate.item())

    def compton(self, ge, ma, ntargets, days_exposure, threshold):
        return self._scatter(0, P.icompton_consts(ma, ge, self.det_z), ntargets, days_exposure, threshold)


class PhotonEventGenerator(_EventGenerator):
    """ALP-photon coupling: decay and inverse-Primakoff rates, a -> gamma gamma decay MC (generators.py:57-128)."""

    def propagate_isotropic(self, new_gagamma=1.0):
        pass

    def inverse_primakoff(self, gagamma, ma, ntargets, days_exposure, threshold=0.0):
        return self._scatter(1, P.iprimakoff_consts(ma, gagamma, self.det_z), ntargets, days_exposure, threshold)

    def simulate_decay_4vectors(self, days_exposure, n_samples=1, uniforms=None, seed=None, parent_phis=None,
                                dtype=np.float64, strict=None):
        """2-body decay a -> gamma gamma of every flux sample, `n_samples` decays each (generators.py:94-128).

        Returns (photon-1 4-vectors, photon-2 4-vectors, weights): two LorentzVectorArray (sequence of
        LorentzVector) and a float64 array, length n_flux*n_samples in parent-major order.  Parents are forward
        (theta = 0) when `flux.axion_angle` is empty (generators.py:111-112); otherwise each parent points along
        (theta_i, phi_i) with phi_i drawn like the reference draws it (or given as `parent_phis`).
        `uniforms` (parity mode): array [n_flux, 2, n_samples] -- per parent the phi draws then the theta draws.
        `dtype=np.float32`: optional single-precision mode (forward parents, free-running RNG): float 4-vectors and
        weights, 36 B/event instead of 72, relative accuracy ~1e-7 of the parent energy (north star bound 1e-5).
        `strict=True` (default `flux.strict` / config.STRICT): return the reference's containers -- two Python lists of
        `LorentzVector` objects and a list of weights (generators.py:124-128); meant for small n (one Python object per
        event)."""
        fl = self.flux
        fl._ensure_production()
        fl._ensure_propagated()
        rt = fl._rt
        n_flux = 0 if fl._F is None else int(fl._F.shape[0])        # generators.py:98
        if fl._D32 is None or int(fl._D32.shape[0]) != n_flux:
            raise ValueError("propagate() must run before simulate_decay_4vectors()")
        self._need_f32(fl._D32, "decay_axion_weight")
        ns = int(n_samples)
        n = n_flux * ns
        ev = P.event_params(days_exposure, 0.0)
        angled = len(fl.axion_angle) != 0                           # generators.py:111-114
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64).ravel()
            if uniforms.shape[0] != 2 * n:
                raise ValueError(f"uniforms must hold 2*n_flux*n_samples = {2 * n} variates")
            seed = 0
        elif seed is None:
            seed = fl.seed if fl.seed is not None else _draw_seed()
        f32 = np.dtype(dtype) == np.float32
        if f32 and (angled or uniforms is not None):
            raise ValueError("dtype=float32 is available for forward parents with the free-running RNG")
        with rt.activate():
            tdt = torch.float32 if f32 else torch.float64
            p1 = torch.empty((4, n), dtype=tdt, device=rt.device)
            p2 = torch.empty((4, n), dtype=tdt, device=rt.device)
            w = rt.empty(n, tdt)
            d_u = rt.upload(uniforms) if (uniforms is not None and n) else None
            if n and f32:
                table = rt.empty(n_flux * N.PARENT_NP)
                N.call("alpk_decay_parent_rows", N.ptr(fl._E), N.ptr(fl._D32), n_flux, float(fl.ma**2), C.byref(ev), ns,
                       N.ptr(table), rt.stream())
                N.call("alpk_decay2body_f32", N.ptr(table), n_flux, ns, None, C.c_uint64(seed), N.ptr(p1), N.ptr(p2),
                       N.ptr(w), rt.stream())
            elif n and not angled:
                table = rt.empty(n_flux * N.PARENT_NP)
                N.call("alpk_decay_parent_rows", N.ptr(fl._E), N.ptr(fl._D32), n_flux, float(fl.ma**2), C.byref(ev), ns,
                       N.ptr(table), rt.stream())
                N.call("alpk_decay2body", N.ptr(table), n_flux, ns, None, N.ptr(d_u), C.c_uint64(seed),
                       1 if fl.precision != "faithful" else 0, N.ptr(p1), N.ptr(p2), N.ptr(w), rt.stream())
            elif n:
                # parents with a polar angle: 4-vectors assembled per parent exactly as generators.py:117-120 (the
                # azimuths come from the global NumPy stream like the reference's, or from `parent_phis`), then the
                # general decay kernel with the dense 4x4 boost
                ea = np.asarray(fl.axion_energy, dtype=np.float64)
                thetas = np.asarray(fl.axion_angle, dtype=np.float64)
                phis = np.random.uniform(0.0, 2*np.pi, n_flux) if parent_phis is None else np.asarray(parent_phis, dtype=np.float64)
                pa = np.sqrt(ea**2 - fl.ma**2)
                p4 = np.stack([ea, pa*np.cos(phis)*np.sin(thetas), pa*np.sin(phis)*np.sin(thetas), pa*np.cos(thetas)], axis=1)
                ew = days_exposure * P.S_PER_DAY * np.asarray(fl.decay_axion_weight) / ns      # generators.py:97 (float32)
                d_p4, d_w = rt.upload(p4.ravel()), rt.upload(ew.astype(np.float64))
                table = rt.empty(n_flux * N.PARENT_GENERAL_NP)
                N.call("alpk_decay2body_general", N.ptr(d_p4), N.ptr(d_w), n_flux, 0.0, 0.0, ns, None, N.ptr(d_u),
                       C.c_uint64(seed), N.ptr(table), N.ptr(p1), N.ptr(p2), N.ptr(w), None, rt.stream())
        lv1, lv2, wts = LorentzVectorArray(p1), LorentzVectorArray(p2), _DeviceWeights(w)
        if fl.strict if strict is None else strict:
            return list(lv1), list(lv2), list(wts)
        return lv1, lv2, wts


class _DeviceWeights:
    """Event weights of the decay MC: a lazily-copied float64 host array with the device tensor alongside."""

    def __init__(self, t):
        self.device = t
        self._h = None

    def _host(self):
        if self._h is None:
            from .runtime import Runtime
            self._h = Runtime.get(self.device.device).to_host(self.device)
        return self._h

    def __len__(self):
        return int(self.device.shape[0])

    def __getitem__(self, i):
        return self._host()[i]

    def __iter__(self):
        return iter(self._host())

    def __array__(self, dtype=None, copy=None):
        return self._host() if dtype is None else self._host().astype(dtype)

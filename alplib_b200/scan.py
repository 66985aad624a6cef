"""Batched (m_a, g) sensitivity scans.

The reference has no scan driver: users loop `Flux*(ma, g0).simulate(); propagate(new_coupling=g); decays()` over a grid
(README.md:199-226, examples/ex4_na64.py:30-40), re-running the whole chain per point.  `scan_rates` evaluates the same
event-rate map in a single kernel launch (grid = mass x sample tile, threads = couplings), sharded over GPUs by mass.
The samples are the ones the per-point API would generate with the same seed, so each map entry equals the per-point
result up to the order of the final sum.
"""
import collections
import ctypes as C

import numpy as np
import torch
from numpy import power

from . import _native as N
from . import config
from . import params as P
from .constants import ALPHA, M_E, pi
from .decay import W_ee, W_gg, W_gg_loop, b1
from .dist import world as _world
from .fluxes import _as_flux_table, _draw_seed
from .materials import Material
from .photon_xs import AbsCrossSection
from .runtime import Runtime

COMPTON, PRIMAKOFF = 0, 1
_PROC = {"compton": COMPTON, "primakoff": PRIMAKOFF}


_PARTIALS_BYTES = 256 << 20          # cap of the per-call partial-sum scratch (the tile axis is walked in chunks beyond it)
_grid_cache = collections.OrderedDict()     # (device, process, grids, g0, loop_decay) -> device tables of the grid
_flux_cache = collections.OrderedDict()     # (device, table bytes) -> device columns of the flux table
_CACHE_SLOTS = 8


def _cached(cache, key, build):
    hit = cache.get(key)
    if hit is not None:
        cache.move_to_end(key)
        return hit
    val = cache[key] = build()
    while len(cache) > _CACHE_SLOTS:
        cache.popitem(last=False)
    return val


def _grid_tables(rt, proc, ma, g, g0, loop_decay):
    """Device tables that depend only on the grid: ma, ma**2, width[n_ma][n_g], rescale[n_g] -- kept resident across calls
    (a scan is usually repeated on the same grid for several fluxes / exposures / detectors)."""
    def build():
        # decay widths of every grid point, with the operation order of decay.py:10-13, 26-29, 48-52 (vectorised: 4e4 scalar
        # calls would cost more than the scan kernel itself); per-axis powers as Python-float pow (what the scalar width
        # functions evaluate), products broadcast in their order
        g2 = np.array([float(x) ** 2 for x in g])
        ma3 = np.array([float(m) ** 3 for m in ma])
        with np.errstate(all="ignore"):
            if proc == COMPTON:
                open_ = 1 - 4 * (M_E / ma) ** 2 > 0
                root = np.sqrt(np.where(open_, 1 - (2 * M_E / ma) ** 2, 0.0))
                width = np.where(open_[:, None], (g2[None, :] * ma[:, None]) * root[:, None] / (8 * pi), 0.0)
                if loop_decay:      # rare option: keep the scalar functions (exact same widths as the per-point API)
                    width = np.array([[W_ee(gj, m) + W_gg_loop(gj, m, M_E) for gj in g] for m in ma])
            else:
                width = g2[None, :] * ma3[:, None] / (64 * pi)
        rescale = power(g / g0, 2)
        # ONE pinned H2D copy for the four tables
        return tuple(rt.upload_packed([(ma, np.float64), (ma ** 2, np.float64), (np.ascontiguousarray(width).ravel(), np.float64),
                                       (rescale, np.float64)]))
    return _cached(_grid_cache, (rt.index, proc, ma.tobytes(), g.tobytes(), float(g0), bool(loop_decay)), build)


def _flux_columns(rt, pf):
    return _cached(_flux_cache, (rt.index, pf.shape, pf.tobytes()),
                   lambda: tuple(rt.upload_packed([(pf[:, 0], np.float64), (pf[:, 1], np.float64)])))


def scan_rates(process, photon_flux, target, ma_grid, g_grid, n_samples=100, det_dist=4.0, det_length=0.2,
               det_area=0.04, days_exposure=100, threshold=0.0, g0=None, is_isotropic=True, loop_decay=False,
               seed=None, precision=None, device=None, row_offset=0, return_device=False):
    """Decay-event rates on the (m_a, g) grid for `process` ("compton": FluxComptonIsotropic + ElectronEventGenerator;
    "primakoff": FluxPrimakoffIsotropic + PhotonEventGenerator).  Returns rates[n_ma, n_g]: entry (i, j) is what

        f = Flux*Isotropic(photon_flux, target, ..., axion_mass=ma[i], axion_coupling=g0, n_samples=n_samples, seed=seed)
        f.simulate(); f.propagate(new_coupling=g[j]); EventGenerator(f, det).decays(days_exposure, threshold)

    returns.  g0 defaults to g_grid[0]."""
    proc = _PROC[process] if isinstance(process, str) else int(process)
    pf = _as_flux_table(photon_flux, "photon_flux")
    ma = np.ascontiguousarray(ma_grid, dtype=np.float64).ravel()
    g = np.ascontiguousarray(g_grid, dtype=np.float64).ravel()
    n_ma, n_g = int(ma.shape[0]), int(g.shape[0])
    g0 = float(g[0]) if g0 is None else float(g0)
    precision = precision if precision is not None else config.PRECISION
    seed = seed if seed is not None else _draw_seed()
    z = target.z[0]
    if proc == COMPTON:
        c0, c1 = float(g0 ** 2 / 4 / pi), 0.0
        np_row = N.COMPTON_NP
    else:
        c0, c1 = P.primakoff_consts(g0, z)
        np_row = 2
        n_samples = 1
    geom = det_area / (4 * pi * det_dist ** 2) if is_isotropic else None
    prop = P.prop_params(1.0, 0.0, 1.0, det_dist, det_length, geom_accept=geom,
                         fast="fp32" if precision == "fp32" else precision == "fast")
    ev = P.event_params(days_exposure, threshold)
    rt = Runtime.get(device)
    xs = AbsCrossSection(target)
    n_rows = int(pf.shape[0])
    with rt.activate():
        out = rt.empty(n_ma * n_g)
        if n_ma and n_g:
            le, lx, nt = xs.device_table(rt)
            d_e, d_p = _flux_columns(rt, pf)
            d_ma, d_ma2, d_w, d_r = _grid_tables(rt, proc, ma, g, g0, loop_decay)
            n_tiles = max(1, int(N.lib().alpk_scan_tiles(n_rows, int(n_samples))))
            # partial sums [n_ma][tiles of one chunk][n_g]: bounded, the launcher walks the tile axis in chunks
            chunk_tiles = max(1, min(n_tiles, _PARTIALS_BYTES // (8 * n_ma * n_g)))
            tables = rt.empty(max(1, n_ma * n_rows * np_row))
            partials = rt.empty(n_ma * chunk_tiles * n_g)
            N.call("alpk_scan", proc, N.ptr(d_e), N.ptr(d_p), n_rows, int(row_offset), N.ptr(le), N.ptr(lx), nt,
                   N.ptr(d_ma), N.ptr(d_ma2), n_ma, N.ptr(d_w), N.ptr(d_r), n_g, c0, c1, float(z), int(n_samples),
                   C.c_uint64(seed), int(config.PHILOX_ROUNDS[precision]), C.byref(prop), C.byref(ev), N.ptr(tables),
                   N.ptr(partials), chunk_tiles, N.ptr(out),
                   rt.stream())
    out = out.view(n_ma, n_g)
    return out if return_device else rt.to_host(out)


def scan_rates_distributed(process, photon_flux, target, ma_grid, g_grid, group=None, **kw):
    """`scan_rates` with the mass axis sharded over the ranks of a torch.distributed process group (one process per
    GPU).  Masses are dealt round-robin (rank r takes ma[r::world]): the cost of a mass row is very uneven -- stable
    ALPs (m_a < 2 m_e for the electron coupling) are skipped by the kernel, heavy ones have few rows above threshold --
    and contiguous blocks would leave some ranks idle.  Every rank returns the full map; the only communication is one
    all-gather of the ranks' disjoint rows: on one node a single kernel per rank that pushes its rows into every rank's peer
    window over NVLink, already interleaved (peer.PeerGroup.all_gather_rows); otherwise one NCCL all_gather_into_tensor."""
    rank, nranks = _world(group)
    ma = np.ascontiguousarray(ma_grid, dtype=np.float64).ravel()
    g = np.ascontiguousarray(g_grid, dtype=np.float64).ravel()
    kw = dict(kw)
    kw.setdefault("g0", float(g[0]))
    if kw.get("seed") is None and nranks > 1:
        # one Philox key for the whole map: drawn on rank 0, broadcast (each rank drawing its own would make the map depend
        # on the number of ranks)
        import torch.distributed as dist
        box = [_draw_seed() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        kw["seed"] = int(box[0])
    if nranks == 1:
        return scan_rates(process, photon_flux, target, ma, g, **kw)
    import torch.distributed as dist
    kw["return_device"] = True
    n_ma, n_g = int(ma.shape[0]), int(g.shape[0])
    local = scan_rates(process, photon_flux, target, ma[rank::nranks], g, **kw)
    rt = Runtime.get(local.device)
    from .peer import peer_group
    pg = peer_group(group, local)
    if pg is not None and pg.fits_gather(n_ma * n_g):
        with rt.activate():
            full = pg.all_gather_rows(local, rank, nranks, n_ma)
        return rt.to_host(full)
    # every rank's block padded to the same number of rows -> ONE all_gather_into_tensor (a single NCCL call, no per-rank
    # output list), one pinned D2H of the gathered blocks; the round-robin interleave is undone on the host (a reshape)
    n_loc = (n_ma + nranks - 1) // nranks
    if local.shape[0] < n_loc:
        pad = torch.zeros((n_loc, n_g), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        local = pad
    flat = torch.empty((nranks * n_loc, n_g), dtype=local.dtype, device=local.device)      # rank-major concatenation
    if dist.get_backend(group) == "gloo":                       # (CPU tests / more ranks than GPUs: host-staged)
        h = flat.cpu()
        dist.all_gather_into_tensor(h, local.cpu().contiguous(), group=group)
        blocks = h.numpy()
    else:
        dist.all_gather_into_tensor(flat, local.contiguous(), group=group)
        blocks = rt.to_host(flat)
    return np.ascontiguousarray(blocks.reshape(nranks, n_loc, n_g).transpose(1, 0, 2).reshape(nranks * n_loc, n_g)[:n_ma])

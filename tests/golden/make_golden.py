#!/usr/bin/env python
"""
Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference (athompson-git/alplib,
mounted read-only at /root/reference) in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz + kat.json

The reference is imported as package `alplib` through a symlink in a temp dir; `matplotlib` and `vegas` (absent
here, imported at det_xs.py:9 / cross_section_mc.py:9 but unused on this path) are stubbed.  For every case the
script seeds the reference's global RNG with `np.random.seed(s)`, replays the identical uniforms from
`np.random.RandomState(s)` through the oracle (oracle/alp_oracle.py), asserts oracle == reference, and stores
inputs + the injected uniforms + the REFERENCE's outputs.  The GPU box has no /root/reference: the tests only read
the committed fixtures.

"f64" outputs are the reference run with its float32 cast bypassed (np.float32 temporarily aliased to np.float64
while `propagate` executes) -- the quantity the 1e-9 FP64 criterion of the north star is evaluated on.
"""
import contextlib
import json
import os
import sys
import tempfile
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def import_reference():
    tmp = tempfile.mkdtemp(prefix="alplib_ref_")
    os.symlink("/root/reference", os.path.join(tmp, "alplib"))
    sys.path.insert(0, tmp)
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    plt.hist2d = lambda *a, **k: None
    mpl.pyplot = plt
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "vegas": types.ModuleType("vegas")})
    import alplib.fluxes as F
    import alplib.generators as G
    return F, G


F, G = import_reference()
from oracle import alp_oracle as O   # noqa: E402


@contextlib.contextmanager
def f32_cast_bypassed():
    real = np.float32
    np.float32 = np.float64
    try:
        yield
    finally:
        np.float32 = real


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return np.inf
    if a.size == 0:
        return 0.0
    d = np.abs(a - b)
    s = np.maximum(np.abs(b), 1e-300)
    m = (d != 0)
    return float(np.max(d[m] / s[m])) if m.any() else 0.0


def check(name, a, b, tol):
    r = rel(a, b)
    status = "OK " if r <= tol else "FAIL"
    print(f"  [{status}] {name:48s} max rel {r:.3e} (tol {tol:g})")
    if r > tol:
        raise SystemExit(f"oracle disagrees with the reference on {name}")


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"  wrote {os.path.relpath(path, ROOT)} ({os.path.getsize(path)} B)")


def propagate_both(flux, new_coupling=None, **kw):
    """Run the reference propagate twice: stock (float32 attrs) and with the cast bypassed (f64)."""
    args = (new_coupling,) if new_coupling is not None else ()
    with f32_cast_bypassed():
        flux.propagate(*args, **kw)
        d64 = np.array(flux.decay_axion_weight, dtype=np.float64)
        s64 = np.array(flux.scatter_axion_weight, dtype=np.float64)
    flux.propagate(*args, **kw)
    return d64, s64, np.array(flux.decay_axion_weight), np.array(flux.scatter_axion_weight)


GEOM_DEFAULT = dict(det_dist=4.0, det_length=0.2, det_area=0.04)
GEOM_DUNE = dict(det_dist=574.0, det_length=5.0, det_area=21.0)


# -------------------------------------------------------------------------------------------------------------
def kat_scalars():
    print("scalar KATs")
    W = F.Material("W"); C = F.Material("C")
    absW = F.AbsCrossSection(W); absC = F.AbsCrossSection(C)
    k = {
        "METER_BY_MEV": F.METER_BY_MEV, "MEV2_CM2": F.MEV2_CM2, "ALPHA": F.ALPHA, "HBARC": F.HBARC,
        "AVOGADRO": F.AVOGADRO, "M_E": F.M_E, "S_PER_DAY": F.S_PER_DAY, "C_LIGHT": F.C_LIGHT,
        "primakoff_sigma(100,1e-5,0.1,74)": float(F.primakoff_sigma(100, 1e-5, 0.1, 74)),
        "absW.sigma_mev(100)": float(absW.sigma_mev(100)), "absW.sigma_mev(1.0)": float(absW.sigma_mev(1.0)),
        "absW.sigma_mev(2e5)": float(absW.sigma_mev(2e5)),
        "absW.rows": int(absW.pe_data.shape[0]), "absC.rows": int(absC.pe_data.shape[0]),
        "W_gg(1e-5,0.1)": float(F.W_gg(1e-5, 0.1)), "W_ee(1e-6,10)": float(F.W_ee(1e-6, 10)),
        "W_ee(1e-7,1.5)": float(F.W_ee(1e-7, 1.5)), "W_ee(1e-7,0.9)": float(F.W_ee(1e-7, 0.9)),
        "W_gg_loop(1e-6,10,M_E)": float(F.W_gg_loop(1e-6, 10, F.M_E)),
        "W_gg_loop(1e-6,0.5,M_E)": float(F.W_gg_loop(1e-6, 0.5, F.M_E)),
        "compton_dsigma_dea(30,100,1e-5,0.1,74)": float(F.compton_dsigma_dea(30, 100, 1e-5, 0.1, 74)),
        "compton_dsigma_dea(99.99,100,1e-5,0.1,74)": float(F.compton_dsigma_dea(99.99, 100, 1e-5, 0.1, 74)),
        "brem_dsigma_dea(300,1000,1e-6,10,6)": float(F.brem_dsigma_dea(300, 1000, 1e-6, 10, 6)),
        "brem_dsigma_dx_vector(0.3,1e-6,10,6)": float(F.brem_dsigma_dx_vector(0.3, 1e-6, 10, 6)),
        "track_length_integrated_prob(1000,400,5)": float(F.track_length_integrated_prob(1000, 400, 5)),
        "nu_integral(ln2.5,-1,20/3)": float(F.nu_integral(np.log(2.5), -1, 20 / 3)),
        "icompton_sigma([50],0.1,1e-5,18)": float(F.icompton_sigma(np.array([50.0]), 0.1, 1e-5, 18)[0]),
        "iprimakoff_sigma(100,1e-5,0.1,18)": float(F.iprimakoff_sigma(100, 1e-5, 0.1, 18)),
        "W.ntargets": float(W.ntargets), "W.rad_length": float(W.rad_length), "C.rad_length": float(C.rad_length),
    }
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(k, f, indent=1, sort_keys=True)
    # pin the oracle
    oW = O.AbsCrossSection(O.Material("W"))
    check("abs table rows", oW.pe_data.shape[0], k["absW.rows"], 0)
    check("sigma_mev(100)", oW.sigma_mev(100), k["absW.sigma_mev(100)"], 1e-15)
    check("sigma_mev(2e5)", oW.sigma_mev(2e5), k["absW.sigma_mev(2e5)"], 1e-15)
    check("primakoff_sigma", O.primakoff_sigma(100, 1e-5, 0.1, 74), k["primakoff_sigma(100,1e-5,0.1,74)"], 1e-15)
    check("W_gg_loop hi", O.W_gg_loop(1e-6, 10, O.M_E), k["W_gg_loop(1e-6,10,M_E)"], 1e-15)
    check("W_gg_loop lo", O.W_gg_loop(1e-6, 0.5, O.M_E), k["W_gg_loop(1e-6,0.5,M_E)"], 1e-15)
    check("tlip", O.track_length_integrated_prob(1000, 400, 5), k["track_length_integrated_prob(1000,400,5)"], 1e-15)
    # table values on a dense grid (array call) for the device interpolation test
    e = np.concatenate([np.logspace(-3.2, 5.2, 400), absW.pe_data[:, 0]])
    save("abs_table_W", energy=e, sigma_mev=absW.sigma_mev(e), table_e=absW.pe_data[:, 0], table_xs=absW.pe_data[:, 1])
    eC = np.concatenate([np.logspace(-3.2, 5.2, 200), absC.pe_data[:, 0]])
    save("abs_table_C", energy=eC, sigma_mev=absC.sigma_mev(eC), table_e=absC.pe_data[:, 0], table_xs=absC.pe_data[:, 1])


# -------------------------------------------------------------------------------------------------------------
def case_primakoff():
    print("Primakoff (C1 KAT + ragged table)")
    W, Ar = F.Material("W"), F.Material("Ar")
    # C1
    pf = np.array([[100.0, 1e12]])
    fl = F.FluxPrimakoffIsotropic(pf, W, axion_mass=0.1, axion_coupling=1e-5, n_samples=1000, **GEOM_DEFAULT)
    fl.simulate()
    d64, s64, d32, s32 = propagate_both(fl)
    gen = G.PhotonEventGenerator(fl, Ar)
    rate = gen.decays(100, 5.0)
    ipr = gen.inverse_primakoff(1e-5, 0.1, 1e27, 100, 5.0)
    oE, oF = O.primakoff_simulate(pf, 0.1, 1e-5, 74, O.AbsCrossSection(O.Material("W")))
    check("C1 axion_flux", oF, fl.axion_flux, 1e-14)
    op = O.propagate(oE, oF, 0.1, O.W_gg(1e-5, 0.1), 1.0, 4.0, 0.2, geom_accept=O.geom_acceptance(0.04, 4.0))
    check("C1 decay_w f32", op["decay_w"], d32, 0)
    check("C1 decay_w f64", op["decay_w64"] , d64 / np.float64(O.geom_acceptance(0.04, 4.0)), 1e-6)
    _, orate = O.decays(oE, op["decay_w"], 100, 5.0)
    check("C1 decays()", orate, rate, 1e-15)
    save("c1_primakoff", photon_flux=pf, ma=0.1, g=1e-5, axion_energy=fl.axion_energy, axion_flux=fl.axion_flux,
         decay_w32=d32, scatter_w32=s32, decay_w64=d64, scatter_w64=s64, decays_100_5=rate,
         decay_weights=gen.decay_weights, iprimakoff_1e27_100_5=ipr, scatter_weights=gen.scatter_weights)

    # ragged: rows below ma, rows outside the XCOM table, duplicated knots
    rs = np.random.RandomState(42)
    eg = np.concatenate([10 ** rs.uniform(-2.5, 4.5, 300), [0.05, 1.809e-3, 2e5, 1e5, 1e-3, 0.7, 0.7]])
    ph = 10 ** rs.uniform(6, 14, eg.shape[0])
    pf = np.stack([eg, ph], axis=1)
    fl = F.FluxPrimakoffIsotropic(pf, W, axion_mass=0.7, axion_coupling=3e-6, **GEOM_DEFAULT)
    fl.simulate()
    d64, s64, d32, s32 = propagate_both(fl)
    d64b, s64b, d32b, s32b = propagate_both(fl, new_coupling=2e-5)
    gen = G.PhotonEventGenerator(fl, Ar)
    rate = gen.decays(30, 1.0)
    oE, oF = O.primakoff_simulate(pf, 0.7, 3e-6, 74, O.AbsCrossSection(O.Material("W")))
    check("ragged E", oE, fl.axion_energy, 0)
    check("ragged flux", oF, fl.axion_flux, 1e-14)
    ga = O.geom_acceptance(0.04, 4.0)
    op = O.propagate(oE, oF, 0.7, O.W_gg(2e-5, 0.7), np.power(2e-5 / 3e-6, 2), 4.0, 0.2, geom_accept=ga)
    check("ragged decay_w f64 (new coupling)", op["decay_w64"] * ga, d64b, 1e-13)
    check("ragged decay_w f32 (new coupling)", op["decay_w"], d32b, 1.3e-7)
    save("primakoff_ragged", photon_flux=pf, ma=0.7, g=3e-6, axion_energy=fl.axion_energy, axion_flux=fl.axion_flux,
         decay_w32=d32, scatter_w32=s32, decay_w64=d64, scatter_w64=s64, new_coupling=2e-5,
         decay_w32_new=d32b, scatter_w32_new=s32b, decay_w64_new=d64b, scatter_w64_new=s64b, decays_30_1=rate)


def case_compton():
    print("Compton (KAT-C + 60 rows x 64)")
    W, Ar = F.Material("W"), F.Material("Ar")
    oabs = O.AbsCrossSection(O.Material("W"))
    # KAT-C
    pf = np.array([[0.3, 1e9], [20.0, 1e9], [50.0, 2e9]])
    np.random.seed(123)
    fl = F.FluxComptonIsotropic(pf, W, axion_mass=1.5, axion_coupling=1e-7, n_samples=2, **GEOM_DEFAULT)
    fl.simulate()
    d64, s64, d32, s32 = propagate_both(fl)
    gen = G.ElectronEventGenerator(fl, Ar)
    r_dec = gen.decays(100, 5.0)
    r_com = gen.compton(1e-7, 1.5, 1e27, 100, 5.0)
    d64n, s64n, d32n, s32n = propagate_both(fl, new_coupling=3e-7)
    us = O.UniformSource(np.random.RandomState(123))
    oE, oF = O.compton_simulate(pf, 1.5, 1e-7, 74, oabs, 2, us)
    check("KAT-C E", oE, fl.axion_energy, 0)
    check("KAT-C flux", oF, fl.axion_flux, 1e-14)
    save("katc_compton", photon_flux=pf, ma=1.5, g=1e-7, n_samples=2, uniforms=np.concatenate(us.log),
         axion_energy=fl.axion_energy, axion_flux=fl.axion_flux, decay_w32=d32, scatter_w32=s32,
         decay_w64=d64, scatter_w64=s64, decays_100_5=r_dec, compton_1e27_100_5=r_com,
         new_coupling=3e-7, decay_w32_new=d32n, scatter_w32_new=s32n, decay_w64_new=d64n, scatter_w64_new=s64n)

    # bigger, ragged (rows under the s threshold), loop_decay, non-isotropic variant
    rs = np.random.RandomState(7)
    eg = np.sort(10 ** rs.uniform(-0.6, 3.0, 60))
    pf = np.stack([eg, 1e16 * eg ** -1.5 * np.exp(-eg / 150)], axis=1)
    for tag, ma, g, loop, iso in (("compton_mid", 1.5, 1e-7, False, True), ("compton_loop", 0.8, 2e-6, True, False)):
        np.random.seed(2024)
        fl = F.FluxComptonIsotropic(pf, W, axion_mass=ma, axion_coupling=g, n_samples=64, loop_decay=loop,
                                    is_isotropic=iso, **GEOM_DEFAULT)
        fl.simulate()
        d64, s64, d32, s32 = propagate_both(fl)
        gen = G.ElectronEventGenerator(fl, Ar)
        r_dec = gen.decays(100, 5.0); dw = np.array(gen.decay_weights)
        r_com = gen.compton(g, ma, 1e27, 100, 5.0); sw = np.array(gen.scatter_weights)
        r_pair = gen.pair_production(g, 1e27, 100, 1.0); pw = np.array(gen.pair_weights)
        us = O.UniformSource(np.random.RandomState(2024))
        oE, oF = O.compton_simulate(pf, ma, g, 74, oabs, 64, us)
        check(tag + " E", oE, fl.axion_energy, 0)
        check(tag + " flux", oF, fl.axion_flux, 1e-13)
        width = O.W_ee(g, ma) + (O.W_gg_loop(g, ma, O.M_E) if loop else 0.0)
        ga = O.geom_acceptance(0.04, 4.0) if iso else None
        op = O.propagate(oE, oF, ma, width, 1.0, 4.0, 0.2, geom_accept=ga)
        check(tag + " decay_w f64", op["decay_w64"] * (ga if iso else 1.0), d64, 1e-13)
        check(tag + " decay_w f32", op["decay_w"], d32, 1.3e-7)
        _, orate = O.decays(oE, op["decay_w"], 100, 5.0)
        check(tag + " decays()", orate, r_dec, 1e-7)
        _, ocom = O.compton_events(oE, op["scatter_w"], g, ma, 1e27, 0.04, 18, 100, 5.0)
        check(tag + " compton()", ocom, r_com, 1e-7)
        opw, opair = O.pair_production_events(oE, op["scatter_w"], O.ALPPairProdutionCrossSection(O.Material("Ar"), ma), g,
                                              1e27, 0.04, 100, 1.0)
        check(tag + " pair_production()", opair, r_pair, 1e-12)
        save(tag, photon_flux=pf, ma=ma, g=g, n_samples=64, loop_decay=loop, is_isotropic=iso,
             uniforms=np.concatenate(us.log), axion_energy=fl.axion_energy, axion_flux=fl.axion_flux,
             decay_w32=d32, scatter_w32=s32, decay_w64=d64, scatter_w64=s64, width=width,
             decays_100_5=r_dec, compton_1e27_100_5=r_com, decay_weights=dw, scatter_weights=sw,
             pair_1e27_100_1=r_pair, pair_weights=pw)


def case_propagate_misc():
    print("propagate: zero width / time-of-flight cut")
    W = F.Material("W")
    pf = np.array([[5.0, 1e9], [80.0, 3e9], [900.0, 1e8]])
    np.random.seed(5)
    fl = F.FluxComptonIsotropic(pf, W, axion_mass=0.9, axion_coupling=1e-6, n_samples=8, **GEOM_DEFAULT)   # W_ee=0 below 2me
    fl.simulate()
    d64, s64, d32, s32 = propagate_both(fl)
    E = np.array(fl.axion_energy); Fx = np.array(fl.axion_flux)
    op = O.propagate(E, Fx, 0.9, O.W_ee(1e-6, 0.9), 1.0, 4.0, 0.2, geom_accept=O.geom_acceptance(0.04, 4.0))
    check("zero-width decay_w", op["decay_w"], d32, 0)
    check("zero-width scatter_w", op["scatter_w"], s32, 1.3e-7)
    # time-of-flight cut through the base class (fluxes.py:54-58); mutates wgt copy only
    fl2 = F.FluxComptonIsotropic(pf, W, axion_mass=0.9, axion_coupling=1e-6, n_samples=8, **GEOM_DEFAULT)
    fl2.axion_energy = list(E); fl2.axion_flux = list(Fx)
    fl2.timing_window = 2e-10
    with f32_cast_bypassed():
        F.AxionFlux.propagate(fl2, 3e-12, 2.0, time_of_flight_cut=True)
        t64d = np.array(fl2.decay_axion_weight); t64s = np.array(fl2.scatter_axion_weight)
    op2 = O.propagate(E, Fx, 0.9, 3e-12, 2.0, 4.0, 0.2, timing_window=2e-10)
    check("tof decay_w f64", op2["decay_w64"], t64d, 1e-13)
    save("propagate_misc", axion_energy=E, axion_flux=Fx, ma=0.9, zero_decay_w32=d32, zero_scatter_w32=s32,
         tof_width=3e-12, tof_rescale=2.0, tof_window=2e-10, tof_decay_w64=t64d, tof_scatter_w64=t64s)


def case_brem():
    print("Brem (KAT-B + vector + no-track-length + 40 rows x 16)")
    C = F.Material("C")
    oC = O.Material("C")
    ef = np.array([[500.0, 1e12], [2000.0, 5e11]])
    pfx = np.array([[400.0, 3e11], [2500.0, 1e11]])
    np.random.seed(5)
    fl = F.FluxBremIsotropic(ef, pfx, C, target_length=100, axion_mass=10, axion_coupling=1e-6, n_samples=2, **GEOM_DUNE)
    fl.simulate()
    d64, s64, d32, s32 = propagate_both(fl)
    us = O.UniformSource(np.random.RandomState(5))
    oE, oF, info = O.brem_simulate(ef, pfx, 10, 1e-6, oC, 2, us)
    check("KAT-B E", oE, fl.axion_energy, 1e-15)
    check("KAT-B flux", oF, fl.axion_flux, 1e-13)
    save("katb_brem", electron_flux=ef, positron_flux=pfx, ma=10.0, g=1e-6, n_samples=2, target_length=100.0,
         uniforms=np.concatenate(us.log), axion_energy=fl.axion_energy, axion_flux=fl.axion_flux,
         decay_w32=d32, scatter_w32=s32, decay_w64=d64, scatter_w64=s64,
         row_u=info["row_u"], row_E1=info["row_E1"], row_W=info["row_W"], row_emit=info["row_emit"], rows=info["rows"])

    rs = np.random.RandomState(11)
    e = np.linspace(2.0, 3000.0, 40)
    ef = np.stack([e, 1e14 * np.exp(-e / 800)], axis=1)
    pfx = np.stack([e, 0.3e14 * np.exp(-e / 800)], axis=1)
    for tag, ma, boson, tl in (("brem_ps", 10.0, "pseudoscalar", True), ("brem_vec", 3.0, "vector", True),
                               ("brem_notl", 10.0, "pseudoscalar", False)):
        np.random.seed(99)
        fl = F.FluxBremIsotropic(ef, pfx, C, target_length=100, axion_mass=ma, axion_coupling=1e-6, n_samples=16,
                                 boson_type=boson, **GEOM_DUNE)
        fl.simulate(use_track_length=tl)
        d64, s64, d32, s32 = propagate_both(fl)
        us = O.UniformSource(np.random.RandomState(99))
        oE, oF, info = O.brem_simulate(ef, pfx, ma, 1e-6, oC, 16, us, boson_type=boson, use_track_length=tl)
        check(tag + " E", oE, fl.axion_energy, 1e-15)
        check(tag + " flux", oF, fl.axion_flux, 1e-12)
        save(tag, electron_flux=ef, positron_flux=pfx, ma=ma, g=1e-6, n_samples=16, target_length=100.0,
             boson_type=boson, use_track_length=tl, uniforms=np.concatenate(us.log),
             axion_energy=fl.axion_energy, axion_flux=fl.axion_flux, decay_w32=d32, scatter_w32=s32,
             decay_w64=d64, scatter_w64=s64, row_u=info["row_u"], row_E1=info["row_E1"], row_W=info["row_W"],
             row_emit=info["row_emit"], rows=info["rows"])


def case_resonance():
    print("Resonance (KAT-R)")
    C = F.Material("C")
    e = np.linspace(10, 5000, 200)
    pfx = np.stack([e, 3e11 * np.exp(-e / 600)], axis=1)
    np.random.seed(9)
    fl = F.FluxResonanceIsotropic(pfx, C, axion_mass=10, axion_coupling=1e-6, n_samples=1000, **GEOM_DUNE)
    fl.simulate()
    d64, s64, d32, s32 = propagate_both(fl)
    us = O.UniformSource(np.random.RandomState(9))
    oE, oF = O.resonance_simulate(pfx, 10, 1e-6, O.Material("C"), 1000, us)
    check("KAT-R E", oE, fl.axion_energy, 0)
    check("KAT-R flux", oF, fl.axion_flux, 1e-13)
    save("katr_resonance", positron_flux=pfx, ma=10.0, g=1e-6, n_samples=1000, uniforms=np.concatenate(us.log),
         axion_energy=fl.axion_energy, axion_flux=fl.axion_flux, decay_w32=d32, scatter_w32=s32,
         decay_w64=d64, scatter_w64=s64)


def case_pairann():
    print("PairAnnihilation (KAT-A + 12 rows x 8)")
    C = F.Material("C")
    for tag, pfx, ns, seed in (("kata_pairann", np.array([[300.0, 1e11], [900.0, 1e10]]), 2, 11),
                               ("pairann_mid", np.stack([np.linspace(20, 4000, 12), 1e11 * np.exp(-np.linspace(20, 4000, 12) / 900)], axis=1), 8, 21)):
        np.random.seed(seed)
        fl = F.FluxPairAnnihilationIsotropic(pfx, C, axion_mass=10, axion_coupling=1e-6, n_samples=ns, **GEOM_DUNE)
        fl.simulate()
        d64, s64, d32, s32 = propagate_both(fl)
        us = O.UniformSource(np.random.RandomState(seed))
        oE, oF = O.pairann_simulate(pfx, 10, 1e-6, O.Material("C"), ns, us)
        check(tag + " E", oE, fl.axion_energy, 1e-13)
        check(tag + " flux", oF, fl.axion_flux, 1e-12)
        save(tag, positron_flux=pfx, ma=10.0, g=1e-6, n_samples=ns, uniforms=np.concatenate(us.log),
             axion_energy=fl.axion_energy, axion_flux=fl.axion_flux, decay_w32=d32, scatter_w32=s32,
             decay_w64=d64, scatter_w64=s64)


def case_decay4():
    print("simulate_decay_4vectors (KAT-D + boosted + heavy)")
    W, Ar = F.Material("W"), F.Material("Ar")
    for tag, pf, ma, g, ns, seed in (
            ("katd_decay4", np.array([[100.0, 1e12], [40.0, 1e11]]), 0.1, 1e-5, 3, 7),
            ("decay4_boosted", np.array([[1e5, 1e12], [3e4, 1e11], [100.0, 1e11]]), 0.1, 1e-5, 200, 3),
            ("decay4_heavy", np.array([[10.0, 1e12], [5.5, 1e11]]), 5.0, 1e-7, 200, 3)):
        fl = F.FluxPrimakoffIsotropic(pf, W, axion_mass=ma, axion_coupling=g, **GEOM_DEFAULT)
        fl.simulate(); fl.propagate()
        gen = G.PhotonEventGenerator(fl, Ar)
        np.random.seed(seed)
        l1, l2, wl = gen.simulate_decay_4vectors(100, n_samples=ns)
        p1 = np.array([lv.pmu for lv in l1]); p2 = np.array([lv.pmu for lv in l2]); w = np.array(wl)
        us = O.UniformSource(np.random.RandomState(seed))
        o1, o2, ow = O.decay_4vectors(fl.axion_energy, fl.decay_axion_weight, ma, 100, ns, us)
        scale = np.repeat(np.array(fl.axion_energy), ns)[:, None]
        d1 = float(np.max(np.abs(o1 - p1) / scale)); d2 = float(np.max(np.abs(o2 - p2) / scale))
        exact = bool(np.array_equal(o1, p1) and np.array_equal(o2, p2))
        print(f"  {tag}: max |dp|/E_parent = {max(d1, d2):.3e}; bit-identical: {exact}")
        if max(d1, d2) > 1e-13:
            raise SystemExit("oracle disagrees with the reference on " + tag)
        check(tag + " weights", ow, w, 0)
        save(tag, photon_flux=pf, ma=ma, g=g, n_samples=ns, days=100, uniforms=np.concatenate(us.log),
             axion_energy=fl.axion_energy, decay_w32=np.array(fl.decay_axion_weight), p1=p1, p2=p2, w=w)


def case_vol_int():
    print("propagate_iso_vol_int (box / sphere / cylinder)")
    W = F.Material("W")
    pf = np.array([[3.0, 1e9], [40.0, 3e9], [700.0, 1e8]])
    for tag, geo, params, ma, g in (("volint_box", "box", [1.0, 2.0, 0.5], 1.5, 1e-4), ("volint_sphere", "sphere", [0.8], 1.5, 3e-5),
                                    ("volint_cyl", "cylinder", [0.6, 1.2], 2.0, 1e-6)):
        np.random.seed(77)
        fl = F.FluxComptonIsotropic(pf, W, axion_mass=ma, axion_coupling=g, n_samples=12, **GEOM_DEFAULT)
        fl.simulate()
        np.random.seed(78)
        geom = F.DetectorGeometry(4.0, geo, params, n_samples=700)
        with f32_cast_bypassed():
            fl.propagate_iso_vol_int(geom, new_coupling=2 * g)
            d64 = np.array(fl.decay_axion_weight, dtype=np.float64); s64 = np.array(fl.scatter_axion_weight, dtype=np.float64)
        fl.propagate_iso_vol_int(geom, new_coupling=2 * g)
        d32 = np.array(fl.decay_axion_weight); s32 = np.array(fl.scatter_axion_weight)
        us = O.UniformSource(np.random.RandomState(78))
        og = O.DetectorGeometry(4.0, geo, params, 700, us)
        op = O.propagate_iso_vol_int(fl.axion_energy, fl.axion_flux, ma, O.W_ee(2 * g, ma), np.power(2 * g / g, 2), 4.0, og)
        check(tag + " decay_w f64", op["decay_w64"], d64, 1e-13)
        check(tag + " scatter_w f64", op["scatter_w64"], s64, 1e-13)
        check(tag + " decay_w f32", op["decay_w"], d32, 1.3e-7)
        save(tag, photon_flux=pf, ma=ma, g=g, new_coupling=2 * g, geometry=geo, params=np.array(params), l0=4.0, n_geom=700,
             geom_uniforms=np.concatenate(us.log), axion_energy=fl.axion_energy, axion_flux=fl.axion_flux,
             decay_w32=d32, scatter_w32=s32, decay_w64=d64, scatter_w64=s64)


def case_nuclear_pi0():
    print("FluxNuclearIsotropic / FluxPi0Isotropic")
    W = F.Material("W")
    rates = np.array([[0.478, 1e10, 1.0, 0.5], [2.2, 3e9, 0.9, 0.4], [7.6, 1e8, 1.1, 0.6], [0.05, 1e12, 1.0, 0.5]])
    fl = F.FluxNuclearIsotropic(rates, W, axion_mass=0.3, gann0=2e-4, gann1=1e-3, gae=1e-6, **GEOM_DEFAULT)
    fl.simulate(j=1, delta=0.2)
    d64, s64, d32, s32 = propagate_both(fl)
    oE, oF = O.nuclear_simulate(rates, 0.3, 2e-4, 1e-3, 1, 0.2)
    check("nuclear E", oE, fl.axion_energy, 0)
    check("nuclear flux", oF, fl.axion_flux, 1e-15)
    save("nuclear", rates=rates, ma=0.3, gann0=2e-4, gann1=1e-3, gae=1e-6, j=1, delta=0.2, axion_energy=fl.axion_energy,
         axion_flux=fl.axion_flux, decay_w32=d32, scatter_w32=s32, decay_w64=d64, scatter_w64=s64)
    rs = np.random.RandomState(3)
    mom = 10 ** rs.uniform(0.5, 3.5, 400)
    np.random.seed(17)
    fp = F.FluxPi0Isotropic(pi0_rate=0.0259, boson_mass=30.0, coupling=1e-3, det_dist=20.0, det_area=2.0, det_length=1.0, target=W)
    fp.simulate_flux(mom, energy_cut=50.0, angle_cut=0.3)
    fp.propagate()
    us = O.UniformSource(np.random.RandomState(17))
    oE, oF = O.pi0_simulate_flux(mom, 30.0, 1e-3, 0.0259, us, 50.0, 0.3)
    check("pi0 E", oE, fp.axion_energy, 1e-15)
    check("pi0 flux", oF, fp.axion_flux, 1e-15)
    save("pi0_flux", momenta=mom, ma=30.0, g=1e-3, rate=0.0259, energy_cut=50.0, angle_cut=0.3, uniforms=np.concatenate(us.log),
         axion_energy=fp.axion_energy, axion_flux=fp.axion_flux, scatter_w=np.array(fp.scatter_axion_weight),
         decay_w=np.array(fp.decay_axion_weight))
    fr = F.FluxPi0Isotropic(pi0_rate=0.0259, boson_mass=30.0, coupling=1e-3, n_samples=5, target=W)
    fr.simulate(); fr.propagate(is_isotropic=False)
    save("pi0_rest", axion_energy=fr.axion_energy, axion_flux=fr.axion_flux, scatter_w=np.array(fr.scatter_axion_weight))


def case_pairgamma():
    print("FluxPairAnnihilationGamma")
    C = F.Material("C")
    e = np.linspace(5.0, 3000.0, 14)
    pfx = np.stack([e, 1e11 * np.exp(-e / 900)], axis=1)
    np.random.seed(31)
    fl = F.FluxPairAnnihilationGamma(pfx, C, target_radiation_length=42.7, axion_mass=8.0, axion_coupling=1e-5, n_samples=6, **GEOM_DUNE)
    fl.simulate()
    d64, s64, d32, s32 = propagate_both(fl)
    us = O.UniformSource(np.random.RandomState(31))
    oE, oF = O.pairgamma_simulate(pfx, 8.0, 1e-5, O.Material("C"), 6, us, 42.7)
    check("pairgamma E", oE, fl.axion_energy, 1e-15)
    check("pairgamma flux", oF, fl.axion_flux, 1e-13)
    save("pairgamma", positron_flux=pfx, ma=8.0, g=1e-5, n_samples=6, rad_length=42.7, uniforms=np.concatenate([np.atleast_1d(x) for x in us.log]),
         axion_energy=fl.axion_energy, axion_flux=fl.axion_flux, decay_w32=d32, scatter_w32=s32, decay_w64=d64, scatter_w64=s64)


def case_general_decay():
    print("general 2-body decay: neutral-meson flux + simulate_decay_4vectors with polar angles")
    rs = np.random.RandomState(8)
    n = 6
    pz = 10 ** rs.uniform(1.5, 3.5, n); pt = rs.uniform(0, 150, n); ph = rs.uniform(0, 2 * np.pi, n)
    mes = np.stack([pt * np.cos(ph), pt * np.sin(ph), pz, np.sqrt(pz ** 2 + pt ** 2 + F.M_PI0 ** 2)], axis=1)
    for tag, cut, off in (("meson2body", False, 0.0), ("meson2body_cut", True, 0.05)):
        np.random.seed(41)
        fl = F.FluxNeutralMeson2BodyDecay(mes, flux_weight=2.5e3, boson_mass=20.0, coupling=1e-3, meson_mass=F.M_PI0, det_dist=20.0,
                                          det_area=60.0, det_length=1.0, apply_angle_cut=cut, n_samples=40, off_axis_angle=off)
        fl.simulate(); fl.propagate()
        us = O.UniformSource(np.random.RandomState(41))
        oE, oF, oT = O.neutral_meson_simulate(mes, 2.5e3, 20.0, 1e-3, F.M_PI0, 40, 20.0, 60.0, us, cut, off)
        check(tag + " E", oE, fl.axion_energy, 1e-13)
        check(tag + " flux", oF, fl.axion_flux, 1e-14)
        check(tag + " theta", oT, fl.axion_angle, 1e-9)
        save(tag, meson_flux=mes, flux_weight=2.5e3, ma=20.0, g=1e-3, n_samples=40, det_dist=20.0, det_area=60.0, angle_cut=cut,
             off_axis=off, uniforms=np.concatenate(us.log), axion_energy=fl.axion_energy, axion_flux=fl.axion_flux,
             axion_angle=fl.axion_angle, scatter_w=np.array(fl.scatter_axion_weight))
    # decay MC of parents with a polar angle: the dense-boost branch of generators.py:117-121
    W, Ar = F.Material("W"), F.Material("Ar")
    pf = np.array([[300.0, 1e12], [60.0, 1e11], [2000.0, 1e10]])
    fl = F.FluxPrimakoffIsotropic(pf, W, axion_mass=5.0, axion_coupling=1e-6, **GEOM_DEFAULT)
    fl.simulate(); fl.propagate()
    fl.axion_angle = [0.02, 0.3, 0.001]
    gen = G.PhotonEventGenerator(fl, Ar)
    np.random.seed(19)
    l1, l2, wl = gen.simulate_decay_4vectors(100, n_samples=50)
    p1 = np.array([lv.pmu for lv in l1]); p2 = np.array([lv.pmu for lv in l2])
    rs2 = np.random.RandomState(19)
    phis = rs2.uniform(0.0, 2 * np.pi, 3)
    us = O.UniformSource(rs2)
    o1, o2 = [], []
    for i in range(3):
        ea = fl.axion_energy[i]; pa = np.sqrt(ea ** 2 - 5.0 ** 2); th = fl.axion_angle[i]
        a, b = O.decay2body_general((ea, pa * np.cos(phis[i]) * np.sin(th), pa * np.sin(phis[i]) * np.sin(th), pa * np.cos(th)),
                                    0.0, 0.0, 50, us)
        o1.append(a); o2.append(b)
    o1, o2 = np.concatenate(o1), np.concatenate(o2)
    scale = np.repeat(np.array(fl.axion_energy), 50)[:, None]
    d = max(np.max(np.abs(o1 - p1) / scale), np.max(np.abs(o2 - p2) / scale))
    print(f"  decay4_angled: max |dp|/E_parent = {d:.3e}")
    if d > 1e-13:
        raise SystemExit("oracle disagrees with the reference on decay4_angled")
    save("decay4_angled", photon_flux=pf, ma=5.0, g=1e-6, n_samples=50, days=100, axion_angle=np.array(fl.axion_angle),
         parent_phi_uniforms=phis / (2 * np.pi), parent_phis=phis, uniforms=np.concatenate(us.log), axion_energy=fl.axion_energy,
         decay_w32=np.array(fl.decay_axion_weight), p1=p1, p2=p2, w=np.array(wl))


def case_meson3body():
    """Charged-meson 3-body decay fluxes (fluxes.py:605-922): every interaction model of the isotropic class on pion and
    kaon tables, and the decay-in-flight class (exponential decay positions, solid-angle acceptance, energy cut)."""
    print("charged-meson 3-body decays")
    models = ["scalar_ib1", "pseudoscalar_ib1", "vector_ib1", "scalar_ib2", "vector_ib2", "vector_ib9", "sd", "vector_contact"]
    mf = np.array([[0.0, 0.0259], [120.0, 1.0], [900.0, 2.5], [4000.0, 0.3]])
    abd = (1.0, 0.5, 0.2)
    ns = 6
    iso = {}
    k = 0
    for model in models:
        for mt, ma, g in (("pion", 0.1, 1e-3), ("pion", 30.0, 1e-3), ("kaon", 5.0, 1e-2), ("kaon", 120.0, 1.0)):
            seed = 100 + k
            np.random.seed(seed)
            fl = F.FluxChargedMeson3BodyIsotropic(meson_flux=mf, boson_mass=ma, coupling=g, meson_type=mt,
                                                  interaction_model=model, n_samples=ns, c0=-0.97, abd=abd)
            fl.simulate()
            fl.propagate()
            us = O.UniformSource(np.random.RandomState(seed))
            oE, oF = O.charged_meson_isotropic_simulate(mf, ma, g, mt, model, ns, -0.97, [O.M_E, O.M_MU], abd, us)
            check(f"meson3 iso {model} {mt} ma={ma} E", oE, fl.axion_energy, 0.0)
            check(f"meson3 iso {model} {mt} ma={ma} flux", oF, fl.axion_flux, 0.0)
            tag = f"c{k}"
            iso[tag + "_cfg"] = np.array([models.index(model), 0 if mt == "pion" else 1, ma, g])
            iso[tag + "_u"] = np.concatenate(us.log) if us.log else np.zeros(0)
            iso[tag + "_E"] = np.array(fl.axion_energy)
            iso[tag + "_F"] = np.array(fl.axion_flux)
            iso[tag + "_S"] = np.array(fl.scatter_axion_weight)
            k += 1
    save("meson3_iso", meson_flux=mf, abd=np.array(abd), n_samples=ns, c0=-0.97, n_cases=k, **iso)

    mfd = np.array([[300.0, 0.01, 1.0], [800.0, 0.05, 2.0], [3000.0, 0.002, 0.5], [50.0, 0.3, 1.0], [1500.0, 0.02, 0.7]])
    dec = {}
    k = 0
    for model, mt, ma, g, cut in (("scalar_ib1", "pion", 0.1, 1e-3, True), ("scalar_ib1", "kaon", 20.0, 1e-3, False),
                                  ("vector_ib1", "kaon", 0.1, 1.0, True), ("vector_contact", "pion", 20.0, 1e-2, True),
                                  ("pseudoscalar_ib1", "pion", 50.0, 1e-3, True)):
        seed = 300 + k
        np.random.seed(seed)
        fl = F.FluxChargedMeson3BodyDecay(mfd, boson_mass=ma, coupling=g, n_samples=8, meson_type=mt, interaction_model=model,
                                          energy_cut=140.0)
        fl.simulate(cut_on_solid_angle=cut)
        us = O.UniformSource(np.random.RandomState(seed))
        r = O.charged_meson_decay_simulate(mfd, ma, g, mt, model, 8, -0.95, [O.M_E, O.M_MU], 140.0, 541, 10.0, us, cut)
        for key, attr in (("energy", "axion_energy"), ("cosine", "cosines"), ("flux", "axion_flux"), ("angle", "axion_angle"),
                          ("solid_angle", "solid_angles"), ("decay_pos", "decay_pos")):
            check(f"meson3 decay {model} {mt} ma={ma} {key}", r[key], getattr(fl, attr), 0.0)
        # split the replayed stream: one position variate per meson, then 3*n per live (meson, lepton) row
        ulog = np.concatenate(us.log)
        pos_u, blocks, p = [], [], 0
        mm = O.MESON_PARAMS[mt][0]
        n_lep = sum(1 for ml in (O.M_E, O.M_MU) if not (ma > mm - ml))
        for i in range(mfd.shape[0]):
            pos_u.append(ulog[p]); p += 1
            if r["positions"][i] <= 52.0:
                for _ in range(n_lep):
                    blocks.append(ulog[p:p + 24]); p += 24
        assert p == ulog.shape[0]
        with f32_cast_bypassed():
            fl.propagate(decay="electron")
            d64 = np.array(fl.decay_axion_weight, dtype=np.float64); s64 = np.array(fl.scatter_axion_weight, dtype=np.float64)
        fl.propagate(decay="electron")
        tag = f"c{k}"
        dec[tag + "_cfg"] = np.array([models.index(model), 0 if mt == "pion" else 1, ma, g, 1.0 if cut else 0.0])
        dec[tag + "_upos"] = np.array(pos_u)
        dec[tag + "_usamp"] = np.concatenate(blocks) if blocks else np.zeros(0)
        for key in ("energy", "cosine", "flux", "angle", "solid_angle", "decay_pos", "positions"):
            dec[tag + "_" + key] = r[key] if key == "positions" else np.array(getattr(fl, {"energy": "axion_energy", "cosine": "cosines", "flux": "axion_flux", "angle": "axion_angle", "solid_angle": "solid_angles", "decay_pos": "decay_pos"}[key]))
        dec[tag + "_d64"] = d64; dec[tag + "_s64"] = s64
        dec[tag + "_d32"] = np.array(fl.decay_axion_weight); dec[tag + "_s32"] = np.array(fl.scatter_axion_weight)
        k += 1
    save("meson3_decay", meson_flux=mfd, n_samples=8, c0=-0.95, energy_cut=140.0, det_dist=541.0, det_length=10.0,
         n_cases=k, **dec)


def case_meson_mixing():
    """FluxAxionMesonMixing (fluxes.py:1148-1240): pi0 and eta mixing, energy / angle filters, both propagate branches."""
    print("axion-meson mixing")
    rs = np.random.RandomState(77)
    n = 400
    pz = 10 ** rs.uniform(1.5, 4, n)
    pt = pz * 10 ** rs.uniform(-4, -1.2, n)
    phi = rs.uniform(0, 2 * np.pi, n)
    out = {}
    k = 0
    for mtype, mmass, ma, fa in (("Pi0", O.M_PI0, 20.0, 1e5), ("Pi0", O.M_PI0, 300.0, 3e4), ("Eta", O.M_ETA, 700.0, 1e6),
                                 ("Eta", O.M_ETA, 90.0, 2e5)):
        e = np.sqrt(pz ** 2 + pt ** 2 + mmass ** 2) * np.where(rs.random_sample(n) < 0.1, 1e-3, 1.0)   # some below the ALP mass
        mf = np.stack([e, pt * np.cos(phi), pt * np.sin(phi), pz], axis=1)
        fl = F.FluxAxionMesonMixing(meson_flux=mf, meson_mass=mmass, axion_mass=ma, f_a=fa, meson_type=mtype,
                                    det_dist=574.0, det_length=5.0, det_area=21.0)
        fl.simulate()
        oE, oT, oF = O.meson_mixing_simulate(mf, ma, fa, mmass, mtype, 1e21, 2.89, 574.0, 21.0)
        check(f"mixing {mtype} ma={ma} E", oE, fl.axion_energy, 0.0)
        check(f"mixing {mtype} ma={ma} angle", oT, fl.axion_angle, 0.0)
        check(f"mixing {mtype} ma={ma} flux", oF, fl.axion_flux, 0.0)
        assert 0 < len(fl.axion_energy) < n
        assert abs(O.meson_mixing_photon_coupling(ma, fa) / fl.axion_coupling - 1) < 1e-15
        d64, s64, d32, s32 = propagate_both(fl)
        new_fa = 2.5 * fa
        with f32_cast_bypassed():
            fl.propagate(new_fa)
            d64n = np.array(fl.decay_axion_weight, dtype=np.float64); s64n = np.array(fl.scatter_axion_weight, dtype=np.float64)
        assert abs(O.W_agg_hadronic(new_fa, ma) / F.W_agg_hadronic(new_fa, ma) - 1) < 1e-15 if F.W_agg_hadronic(new_fa, ma) else True
        t = f"c{k}"
        out[t + "_cfg"] = np.array([0 if mtype == "Pi0" else 1, mmass, ma, fa, new_fa])
        out[t + "_mf"] = mf
        out[t + "_E"] = np.array(fl.axion_energy); out[t + "_T"] = np.array(fl.axion_angle); out[t + "_F"] = np.array(fl.axion_flux)
        out[t + "_d64"] = d64; out[t + "_s64"] = s64; out[t + "_d32"] = d32; out[t + "_s32"] = s32
        out[t + "_d64n"] = d64n; out[t + "_s64n"] = s64n
        out[t + "_g"] = np.array([fl.axion_coupling, F.W_agg_hadronic(new_fa, ma)])
        k += 1
    save("meson_mixing", n_cases=k, **out)


def case_scatter2to2():
    """Scatter2to2MC (cross_section_mc.py:148-275) with M2InversePrimakoff (README.md:144-158), M2InverseCompton,
    M2Primakoff, M2Compton and M2AssociatedProduction; linear and log sampling; target at rest and a moving second
    particle (boost along a general direction)."""
    print("Scatter2to2MC (README snippet + 2->2 matrix elements)")
    import alplib.cross_section_mc as X
    import alplib.matrix_element as ME
    Ar = F.Material("Ar")
    mN, z = float(Ar.m[0]), Ar.z[0]
    cases = [
        # tag, model, ma, p1 (E, p along z unless given), p2, n, log_sampling, seed
        ("readme_iprim", "inverse_primakoff", 0.1, 100.0, None, 1000, False, 31),
        ("iprim_log", "inverse_primakoff", 0.1, 100.0, None, 64, True, 32),
        ("iprim_heavy", "inverse_primakoff", 25.0, 40.0, None, 64, False, 33),
        ("icompton", "inverse_compton", 0.1, 30.0, None, 64, False, 34),
        ("icompton_log", "inverse_compton", 1.5, 8.0, None, 64, True, 35),
        ("primakoff", "primakoff", 0.5, 50.0, None, 64, False, 36),
        ("compton", "compton", 1.5, 20.0, None, 64, False, 37),
        ("assoc", "associated_production", 10.0, 300.0, None, 64, False, 38),
        ("icompton_moving", "inverse_compton", 0.3, 12.0, (0.8, 0.15, -0.22, 0.31), 64, False, 39),
        ("icompton_below", "inverse_compton", 0.1, 30.0, "below", 8, False, 40),
    ]
    out, kk = {}, 0
    for tag, model, ma, e1, p2spec, n, logs, seed in cases:
        if model == "inverse_primakoff":
            m2r = ME.M2InversePrimakoff(ma=ma, mN=mN, z=z); m_in, m_tgt = ma, mN
        elif model == "primakoff":
            m2r = ME.M2Primakoff(ma=ma, mN=mN, z=z); m_in, m_tgt = 0.0, mN
        elif model == "inverse_compton":
            m2r = ME.M2InverseCompton(ma=ma, z=z); m_in, m_tgt = ma, F.M_E
        elif model == "compton":
            m2r = ME.M2Compton(ma=ma, z=z); m_in, m_tgt = 0.0, F.M_E
        else:
            m2r = ME.M2AssociatedProduction(ma); m_in, m_tgt = F.M_E, F.M_E
        p1 = (e1, 0.0, 0.0, float(np.sqrt(e1 ** 2 - m_in ** 2)))
        if p2spec is None:
            p2 = (m_tgt, 0.0, 0.0, 0.0)
        elif p2spec == "below":                       # s below (m3+m4)^2 is impossible for these elastic-type processes with a
            p2 = (m_tgt, 0.0, 0.0, 0.0)               # physical p1; force it with an off-shell, backward-moving p1 instead
            p1 = (0.2, 0.0, 0.0, -0.19)
            m2r = ME.M2Compton(ma=5.0, z=z); model, ma = "compton", 5.0
        else:
            pe = p2spec
            p2 = (float(np.sqrt(m_tgt ** 2 + pe[1] ** 2 + pe[2] ** 2 + pe[3] ** 2)), pe[1], pe[2], pe[3])
        np.random.seed(seed)
        mc = X.Scatter2to2MC(mtrx2=m2r, p1=X.LorentzVector(*p1), p2=X.LorentzVector(*p2), n_samples=n)
        mc.scatter_sim(log_sampling=logs)
        us = O.UniformSource(np.random.RandomState(seed))
        masses, m2o = O.matrix_element(model, ma, mN=mN, z=z)
        r = O.scatter2to2_sim(masses, m2o, p1, p2, n, us, log_sampling=logs)
        pre = f"c{kk}_"
        out[pre + "model"] = model; out[pre + "ma"] = ma; out[pre + "mN"] = mN; out[pre + "z"] = z
        out[pre + "p1"] = np.array(p1); out[pre + "p2"] = np.array(p2); out[pre + "n"] = n; out[pre + "log"] = logs
        out[pre + "uniforms"] = np.concatenate(us.log)
        if r is None:
            assert len(mc.p3_lab_4vectors) == 0 and mc.dsigma_dcos_cm_wgts.shape[0] == 0, tag
            out[pre + "empty"] = True
            print(f"  [OK ] {tag:48s} below threshold: no samples")
        else:
            cos_r, w_r = mc.get_cosine_lab_weights()
            e3_r, _ = mc.get_e3_lab_weights()
            e4_r, _ = mc.get_e4_lab_weights()
            p3cm = np.array([v.pmu for v in mc.p3_cm_4vectors]); p3lab = np.array([v.pmu for v in mc.p3_lab_4vectors])
            p4lab = np.array([v.pmu for v in mc.p4_lab_4vectors])
            v3lab = np.array([v.vec for v in mc.p3_lab_3vectors]); v4lab = np.array([v.vec for v in mc.p4_lab_3vectors])
            v3cm = np.array([v.vec for v in mc.p3_cm_3vectors])
            scale = abs(p1[0]) + abs(p2[0])
            check(tag + " weights", r["weights"], w_r, 1e-12)
            check(tag + " cosines", r["cosines"], cos_r, 1e-12)
            check(tag + " e3_lab", r["e3_lab"], e3_r, 1e-13)
            check(tag + " p3_lab/scale", r["p3_lab"] / scale + 1, p3lab / scale + 1, 1e-14)
            check(tag + " p4_lab/scale", r["p4_lab"] / scale + 1, p4lab / scale + 1, 1e-14)
            check(tag + " p3_cm/scale", r["p3_cm"] / scale + 1, p3cm / scale + 1, 1e-15)
            check(tag + " v3_lab", r["v3_lab"] + 2, v3lab + 2, 1e-14)
            check(tag + " v4_lab", r["v4_lab"] + 2, v4lab + 2, 1e-12)
            out[pre + "empty"] = False
            out.update({pre + "weights": w_r, pre + "cosines": cos_r, pre + "e3_lab": e3_r, pre + "e4_lab": e4_r,
                        pre + "p3_cm": p3cm, pre + "p3_lab": p3lab, pre + "p4_lab": p4lab, pre + "v3_lab": v3lab,
                        pre + "v4_lab": v4lab, pre + "v3_cm": v3cm,
                        pre + "dsigma_dt_probe": np.array([mc.dsigma_dt(r["s"], t) for t in r["t"][:4]]),
                        pre + "t_probe": r["t"][:4], pre + "s": r["s"],
                        pre + "total_xs": mc.get_total_xs(r["s"]) if model != "associated_production" else np.nan})
        out[pre + "tag"] = tag
        kk += 1
    # matrix elements evaluated directly (the classes are callable: mtrx2(s, t[, coupling_product]))
    sg = np.array([2.0, 50.0, 4.0e3]); tg = -np.array([1e-6, 0.3, 7.0])
    out["m2_s"], out["m2_t"] = sg, tg
    out["m2_inverse_compton"] = np.array([ME.M2InverseCompton(0.4, 18)(s_, tg, 1e-3) for s_ in sg])
    out["m2_compton"] = np.array([ME.M2Compton(0.4, 18)(s_, tg, 1e-3) for s_ in sg])
    out["m2_inverse_primakoff"] = np.array([ME.M2InversePrimakoff(0.4, mN, 18)(s_ + mN ** 2, tg, 1e-3) for s_ in sg])
    out["m2_primakoff"] = np.array([ME.M2Primakoff(0.4, mN, 18)(s_ + mN ** 2, tg, 1e-3) for s_ in sg])
    out["m2_associated"] = np.array([ME.M2AssociatedProduction(0.4)(s_ + 1.5, tg, 1e-3) for s_ in sg])
    out["atomic_ff_q"] = np.array([1e-4, 0.02, 3.0]); out["atomic_ff"] = ME.AtomicElasticFF(18)(out["atomic_ff_q"])
    check("m2_inverse_compton", np.array([O.m2_inverse_compton(s_, tg, 0.4, 18, 1e-3) for s_ in sg]), out["m2_inverse_compton"], 1e-15)
    check("m2_inverse_primakoff", np.array([O.m2_inverse_primakoff(s_ + mN ** 2, tg, 0.4, mN, 18, 1e-3) for s_ in sg]), out["m2_inverse_primakoff"], 1e-15)
    check("atomic_ff", O.atomic_elastic_ff2(out["atomic_ff_q"], 18), out["atomic_ff"], 1e-15)
    save("scatter2to2", n_cases=kk, **out)


def case_xs_methods():
    """Public methods of the tabulated cross-section classes (photon_xs.py:45-52, 117-128)."""
    print("cross-section class methods")
    import alplib.photon_xs as PXS
    e = np.array([0.0009, 0.0015, 0.9, 1.021, 1.03, 5.0, 100.0, 9.9e4, 2e5])
    out = {"energy": e}
    for mat in ("W", "Ar", "NaI"):
        a = PXS.AbsCrossSection(F.Material(mat))
        out[f"abs_{mat}_cm2"] = a.sigma_cm2(e); out[f"abs_{mat}_mev"] = a.sigma_mev(e); out[f"abs_{mat}_mu"] = a.mu(e, 6.3e22)
    for tag, mat, ma in (("ar01", "Ar", 0.1), ("ar095", "Ar", 0.95), ("ge2", "Ge", 2.0)):
        with np.errstate(all="ignore"):
            x = PXS.ALPPairProdutionCrossSection(F.Material(mat), ma=ma)
            out[f"pair_{tag}_cm2"] = x.sigma_cm2(e); out[f"pair_{tag}_mev"] = x.sigma_mev(e); out[f"pair_{tag}_mu"] = x.mu(e, 2.1e22)
    save("xs_methods", **out)


def _fit_cases():
    """Fixed inputs of the limit-helper fixture; shared with tests/test_fit_cpu.py through fit.json (inputs are stored too)."""
    grid = np.logspace(-8, -3, 700)

    def events(g):            # toy signal: rises as g^4, then dies off when the ALP decays before the detector
        return np.array([1e24 * g ** 4 * np.exp(-(g / 3e-5) ** 2), 3e23 * g ** 4 * np.exp(-(g / 1e-5) ** 2)])
    bkg = np.array([40.0, 12.0])
    obs = np.array([44.0, 9.0])
    return grid, events, bkg, obs


def case_fit():
    """Limit helpers (fit.py:12-121): TwoSidedGridSearch, binary_search, cleanLimitData run in the reference."""
    print("fit.py limit helpers")
    import alplib.fit as RF
    from alplib_b200 import fit as MF
    grid, events, bkg, obs = _fit_cases()
    out = {"grid": grid.tolist(), "background": bkg.tolist(), "observations": obs.tolist()}

    def pair(x):
        return [None if v is None else float(v) for v in x]
    tsgs = {
        "signal_only": dict(observations=np.zeros(2)),
        "signal_only_cl": dict(observations=np.zeros(2), delta_chi2=False),
        "with_background": dict(observations=obs, background=bkg),
        "with_background_cl": dict(observations=obs, background=bkg, delta_chi2=False, target_cl=0.95),
    }
    out["TwoSidedGridSearch"] = {}
    for k, kw in tsgs.items():
        r = RF.TwoSidedGridSearch(events, param_grid=grid, **kw)
        m = MF.TwoSidedGridSearch(events, param_grid=grid, **kw)
        assert pair(r) == pair(m), (k, r, m)
        out["TwoSidedGridSearch"][k] = pair(r)
        print(f"  [OK ] TwoSidedGridSearch {k:24s} {pair(r)}")
    r = RF.TwoSidedGridSearch(lambda g: np.zeros(2), np.zeros(2), grid)
    assert r == (None, None) == MF.TwoSidedGridSearch(lambda g: np.zeros(2), np.zeros(2), grid)
    out["TwoSidedGridSearch"]["no_signal"] = pair(r)
    bs = {
        "increasing": (lambda t: t ** 2, 9.0, 0.0, 10.0, dict(tolerance=1e-6)),
        "decreasing": (lambda t: 100 - t, 40.0, 0.0, 100.0, dict(tolerance=1e-6, is_increasing=False)),
        "coarse": (lambda t: np.exp(t), 5.0, -3.0, 4.0, dict(tolerance=0.1)),
        "edge": (lambda t: t, 1e300, 1.0, 1.0 + 4e-16, dict(tolerance=1e-3)),
    }
    out["binary_search"] = {}
    for k, (fn, stop, lo, hi, kw) in bs.items():
        r = RF.binary_search(fn, stop, lo, hi, **kw)
        m = MF.binary_search(fn, stop, lo, hi, **kw)
        assert float(r) == float(m), (k, r, m)
        out["binary_search"][k] = float(r)
        print(f"  [OK ] binary_search {k:29s} {float(r)!r}")
    masses = np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    lower = np.array([1e-6, 2e-6, 5e-5, 4e-6, 6e-6, 9e-6])
    upper = np.array([1e-4, 1e-4, 1e-5, 8e-5, 5e-5, 2e-5])
    out["cleanLimitData"] = {"masses": masses.tolist(), "lower": lower.tolist(), "upper": upper.tolist()}
    for smooth in (False, True):
        rm, rl = RF.cleanLimitData(masses, lower, upper, apply_smoothing=smooth)
        mm, ml = MF.cleanLimitData(masses, lower, upper, apply_smoothing=smooth)
        assert np.array_equal(rm, mm) and np.array_equal(rl, ml), smooth
        out["cleanLimitData"]["smooth" if smooth else "plain"] = {"masses": rm.tolist(), "limits": rl.tolist()}
        print(f"  [OK ] cleanLimitData smoothing={smooth}")
    path = os.path.join(HERE, "fit.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"  wrote {os.path.relpath(path, ROOT)}")


if __name__ == "__main__":
    if len(sys.argv) > 1:                      # regenerate selected cases only: python make_golden.py case_meson3body
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    kat_scalars()
    case_primakoff()
    case_compton()
    case_propagate_misc()
    case_brem()
    case_resonance()
    case_pairann()
    case_decay4()
    case_vol_int()
    case_nuclear_pi0()
    case_pairgamma()
    case_general_decay()
    case_meson3body()
    case_meson_mixing()
    case_scatter2to2()
    case_xs_methods()
    case_fit()
    print("all golden fixtures written; oracle pinned against the reference")

"""CPU tier: the kernels' per-sample arithmetic (alplib_b200/csrc/alp_math*.cuh, compiled for the host by
tests/hostcheck) against the golden fixtures of the reference.  Same header the CUDA kernels include; only libm
differs (glibc here, libdevice on the GPU).  Tolerance: relative 1e-9 (north star, FP64 mode); observed ~1e-15."""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_decay_close, dptr, load_golden, max_rel
from oracle import alp_oracle as O
from oracle import philox as PX
from alplib_b200 import _native as N
from alplib_b200 import params as P

TOL = 1e-9
GA = P.geom_acceptance(0.04, 4.0)


def table(mat="W"):
    a = O.AbsCrossSection(O.Material(mat))
    return np.ascontiguousarray(a.log10_e), np.ascontiguousarray(a.log10_xs)


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10
    r = PX.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = PX.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(v) for v in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = PX.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(v) for v in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_7_rounds_device_code_matches_numpy(hostcheck):
    """Philox4x32-7 (fast-mode production streams, config.PHILOX_ROUNDS): the kernels' C++ and the NumPy restatement are
    independent implementations of the same round function (pinned at 10 rounds by the Random123 vectors above)."""
    n = 4097
    out = np.empty(n)
    for tag, row in ((PX.TAG_COMPTON_EA, 0), (PX.TAG_BREM_EA, 123456), (PX.TAG_PAIRANN, 2 ** 32 - 1)):
        for rounds in (7, 10):
            hostcheck.hc_philox_rounds(C.c_uint64(0xA1B2C3D4E5F60718), C.c_uint32(tag), C.c_uint32(row), C.c_uint64(n), rounds, dptr(out))
            assert np.array_equal(out, PX.uniform_1(0xA1B2C3D4E5F60718, tag, row, n, rounds=rounds))
    a = PX.uniform_1(5, PX.TAG_COMPTON_EA, 0, 1000, rounds=7)
    b = PX.uniform_1(5, PX.TAG_COMPTON_EA, 0, 1000, rounds=10)
    assert not np.array_equal(a, b) and abs(a.mean() - 0.5) < 0.03 and abs(a.var() - 1 / 12) < 0.01
    old = PX.PRODUCTION_ROUNDS
    try:
        PX.PRODUCTION_ROUNDS = 7                      # default rounds: production tags follow the module setting, others stay 10
        assert np.array_equal(PX.uniform_1(5, PX.TAG_COMPTON_EA, 0, 1000), a)
        assert np.array_equal(PX.uniform_1(5, PX.TAG_BREM_ROW, 0, 8), PX.uniform_1(5, PX.TAG_BREM_ROW, 0, 8, rounds=10))
    finally:
        PX.PRODUCTION_ROUNDS = old


def test_philox_device_code_matches_numpy(hostcheck):
    seed = 0xA1B2C3D4_0000_0007
    out = np.empty(1001)
    hostcheck.hc_philox(C.c_uint64(seed), C.c_uint32(PX.TAG_COMPTON_EA), C.c_uint32(12345), C.c_uint64(1001), 0, dptr(out))
    assert np.array_equal(out, PX.uniform_1(seed, PX.TAG_COMPTON_EA, 12345, 1001))
    out2 = np.empty(2 * 500)
    hostcheck.hc_philox(C.c_uint64(seed), C.c_uint32(PX.TAG_DECAY2), C.c_uint32(3), C.c_uint64(500), 1, dptr(out2))
    u1, u2 = PX.uniform_2(seed, PX.TAG_DECAY2, 3, 500)
    assert np.array_equal(out2[0::2], u1) and np.array_equal(out2[1::2], u2)
    assert 0.0 <= out.min() and out.max() < 1.0 and abs(out.mean() - 0.5) < 0.05


@pytest.mark.parametrize("mat", ["W", "C"])
def test_abs_sigma(hostcheck, mat):
    g = load_golden("abs_table_" + mat)
    le, lx = table(mat)
    e = np.ascontiguousarray(g["energy"])
    out = np.empty_like(e)
    hostcheck.hc_abs_sigma(dptr(le), dptr(lx), le.shape[0], dptr(e), C.c_uint64(e.shape[0]), dptr(out))
    assert max_rel(out, g["sigma_mev"]) < 1e-13
    assert out[e > 1e5 * 1.0000001].max() == pytest.approx(2.5680599121454694e+21)      # out-of-range quirk kept


def run_primakoff(hostcheck, pf, ma, g):
    keep = ~(pf[:, 0] < ma)
    eg = np.ascontiguousarray(pf[keep, 0]); ph = np.ascontiguousarray(pf[keep, 1])
    le, lx = table("W")
    pref, r02 = P.primakoff_consts(g, np.int64(74))
    oe, of = np.empty_like(eg), np.empty_like(eg)
    hostcheck.hc_primakoff(dptr(eg), dptr(ph), C.c_uint64(eg.shape[0]), dptr(le), dptr(lx), le.shape[0],
                           C.c_double(ma), C.c_double(pref), C.c_double(r02), dptr(oe), dptr(of))
    return oe, of


def run_propagate(hostcheck, E, F, prop):
    n = E.shape[0]
    d32, s32 = np.empty(n, np.float32), np.empty(n, np.float32)
    d64, s64 = np.empty(n), np.empty(n)
    hostcheck.hc_propagate(dptr(np.ascontiguousarray(E)), dptr(np.ascontiguousarray(F)), C.c_uint64(n), C.byref(prop),
                           dptr(d32), dptr(s32), dptr(d64), dptr(s64))
    return d32, s32, d64, s64


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("name", ["c1_primakoff", "primakoff_ragged"])
def test_primakoff_chain(hostcheck, name, fast):
    g = load_golden(name)
    ma, gg = float(g["ma"]), float(g["g"])
    E, F = run_primakoff(hostcheck, g["photon_flux"], ma, gg)
    assert np.array_equal(E, g["axion_energy"])
    assert max_rel(F, g["axion_flux"]) < 1e-13
    prop = P.prop_params(ma, O.W_gg(gg, ma), 1.0, 4.0, 0.2, geom_accept=GA, fast=fast)
    d32, s32, d64, s64 = run_propagate(hostcheck, E, F, prop)
    assert max_rel(s64 * GA, g["scatter_w64"]) < TOL
    assert_decay_close(d64 * GA, g["decay_w64"], g["scatter_w64"], TOL)
    assert max_rel(s32, g["scatter_w32"]) <= 1.2e-7
    assert_decay_close(d32, g["decay_w32"], g["scatter_w32"], TOL, extra_rel=1.2e-7)
    days, thr, key = (100, 5.0, "decays_100_5") if name == "c1_primakoff" else (30, 1.0, "decays_30_1")
    ev = P.event_params(days, thr)
    out = np.empty(E.shape[0])
    # (the ragged fixture took decays() after propagate(new_coupling): its rate belongs to decay_w32_new)
    w32 = g["decay_w32" if name == "c1_primakoff" else "decay_w32_new"].astype(np.float32)
    hostcheck.hc_decays(dptr(E), dptr(w32), C.c_uint64(E.shape[0]), C.byref(ev), dptr(out))
    assert max_rel(out.sum(), g[key]) < 1e-12
    if name == "c1_primakoff":
        assert out.sum() == 284.1535949707031
        xs = P.iprimakoff_consts(0.1, 1e-5, np.int64(18))
        hostcheck.hc_scatter(1, dptr(E), dptr(g["scatter_w32"].astype(np.float32)), C.c_uint64(E.shape[0]), xs,
                             C.c_double(1e27 / 0.04), C.c_double(P.METER_BY_MEV ** 2), C.byref(ev), dptr(out))
        assert max_rel(out, g["scatter_weights"]) < 1e-13


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("name", ["katc_compton", "compton_mid", "compton_loop"])
def test_compton_chain(hostcheck, name, fast):
    g = load_golden(name)
    ma, gg, ns = float(g["ma"]), float(g["g"]), int(g["n_samples"])
    pf = g["photon_flux"]
    keep = O.compton_row_valid(pf, ma)
    eg = np.ascontiguousarray(pf[keep, 0]); ph = np.ascontiguousarray(pf[keep, 1])
    le, lx = table("W")
    par = P.compton_params(ma, gg, np.int64(74), ns, fast=fast)
    n = eg.shape[0] * ns
    u = np.ascontiguousarray(g["uniforms"])
    assert u.shape[0] == n
    oe, of = np.empty(n), np.empty(n)
    hostcheck.hc_compton(dptr(eg), dptr(ph), C.c_uint64(eg.shape[0]), C.c_uint32(ns), dptr(le), dptr(lx), le.shape[0],
                         C.byref(par), dptr(u), dptr(oe), dptr(of))
    assert np.array_equal(oe, g["axion_energy"])                  # E_a = ma + (eg-ma)*u is bit exact
    assert max_rel(of, g["axion_flux"]) < (TOL if fast else 1e-12)
    loop = bool(g["loop_decay"]) if "loop_decay" in g else False
    iso = bool(g["is_isotropic"]) if "is_isotropic" in g else True
    width = O.W_ee(gg, ma) + (O.W_gg_loop(gg, ma, O.M_E) if loop else 0.0)
    prop = P.prop_params(ma, width, 1.0, 4.0, 0.2, geom_accept=GA if iso else None, fast=fast)
    d32, s32, d64, s64 = run_propagate(hostcheck, oe, of, prop)
    ga = GA if iso else 1.0
    assert max_rel(s64 * ga, g["scatter_w64"]) < TOL
    assert_decay_close(d64 * ga, g["decay_w64"], g["scatter_w64"], TOL)       # conditioned on 1-exp(-x), see conftest
    assert max_rel(s32, g["scatter_w32"]) <= 1.2e-7
    assert_decay_close(d32, g["decay_w32"], g["scatter_w32"], TOL, extra_rel=1.2e-7)
    if "scatter_weights" in g:
        ev = P.event_params(100, 5.0)
        out = np.empty(n)
        xs = P.icompton_consts(ma, gg, np.int64(18))
        hostcheck.hc_scatter(2 if fast else 0, dptr(oe), dptr(g["scatter_w32"].astype(np.float32)), C.c_uint64(n), xs,
                             C.c_double(1e27 / 0.04), C.c_double(P.METER_BY_MEV ** 2), C.byref(ev), dptr(out))
        assert max_rel(out, g["scatter_weights"], floor=1e-300) < (TOL if fast else 1e-12)
        hostcheck.hc_decays(dptr(oe), dptr(g["decay_w32"].astype(np.float32)), C.c_uint64(n), C.byref(ev), dptr(out))
        assert np.array_equal(out, g["decay_weights"])
    if "new_coupling" in g:
        nc = float(g["new_coupling"])
        prop = P.prop_params(ma, O.W_ee(nc, ma), np.power(nc / gg, 2), 4.0, 0.2, geom_accept=GA, fast=fast)
        d32, s32, d64, s64 = run_propagate(hostcheck, oe, of, prop)
        assert_decay_close(d64 * GA, g["decay_w64_new"], g["scatter_w64_new"], TOL)
        assert_decay_close(d32, g["decay_w32_new"], g["scatter_w32_new"], TOL, extra_rel=1.2e-7)


@pytest.mark.parametrize("fast", [False, True])
def test_propagate_misc(hostcheck, fast):
    g = load_golden("propagate_misc")
    E, F = g["axion_energy"], g["axion_flux"]
    d32, s32, d64, s64 = run_propagate(hostcheck, E, F, P.prop_params(0.9, 0.0, 1.0, 4.0, 0.2, geom_accept=GA, fast=fast))
    assert np.all(d32 == 0) and max_rel(s32, g["zero_scatter_w32"]) <= 1.2e-7
    prop = P.prop_params(0.9, float(g["tof_width"]), float(g["tof_rescale"]), 4.0, 0.2, timing_window=float(g["tof_window"]), fast=fast)
    d32, s32, d64, s64 = run_propagate(hostcheck, E, F, prop)
    assert max_rel(s64, g["tof_scatter_w64"]) < TOL
    assert_decay_close(d64, g["tof_decay_w64"], g["tof_scatter_w64"], TOL)


@pytest.mark.parametrize("fast", [0, 1])
@pytest.mark.parametrize("name", ["katd_decay4", "decay4_boosted", "decay4_heavy"])
def test_decay2body(hostcheck, name, fast):
    g = load_golden(name)
    ma, ns = float(g["ma"]), int(g["n_samples"])
    E = np.ascontiguousarray(g["axion_energy"]); w32 = np.ascontiguousarray(g["decay_w32"].astype(np.float32))
    npar = E.shape[0]
    u = np.ascontiguousarray(g["uniforms"][npar:])              # the first n_flux draws are the unused parent phis
    n = npar * ns
    p1, p2, w = np.empty((4, n)), np.empty((4, n)), np.empty(n)
    ev = P.event_params(int(g["days"]), 0.0)
    hostcheck.hc_decay2body(dptr(E), dptr(w32), C.c_uint64(npar), C.c_double(ma ** 2), C.byref(ev), C.c_uint32(ns),
                            dptr(u), fast, dptr(p1), dptr(p2), dptr(w))
    scale = np.repeat(E, ns)[None, :]
    assert np.max(np.abs(p1 - g["p1"].T) / scale) < TOL          # |dp| / E_parent (SURVEY section 7 hard part 1)
    assert np.max(np.abs(p2 - g["p2"].T) / scale) < TOL
    assert np.array_equal(w, g["w"])
    # the host build follows the reference's op order; what is left is glibc-scalar vs NumPy-SIMD sin/cos/acos
    assert np.max(np.abs(p1 - g["p1"].T) / scale) < 1e-13 and np.max(np.abs(p2 - g["p2"].T) / scale) < 1e-13


# ---- e+- channels ---------------------------------------------------------------------------------------------------
def test_nu_integral_gauss_legendre(hostcheck):
    hostcheck.hc_nu_integral.restype = C.c_double
    gx, gw = P.gauss_legendre()
    for x in (1e-8, 1e-6, 1e-3, 0.1, 0.9162907318741551, 2.0, 9.0, 13.0):
        got = hostcheck.hc_nu_integral(C.c_double(x), C.c_double(20 / 3), dptr(gx), dptr(gw))
        assert got == pytest.approx(O.nu_integral(x, -1, 20 / 3), rel=1e-13), x
    assert hostcheck.hc_nu_integral(C.c_double(np.log(2.5)), C.c_double(20 / 3), dptr(gx), dptr(gw)) == \
        pytest.approx(2.5972770215450707, rel=1e-13)                       # SURVEY section 8c KAT
    assert hostcheck.hc_nu_integral(C.c_double(0.7), C.c_double(0.8), dptr(gx), dptr(gw)) == \
        pytest.approx(O.nu_integral(0.7, -1, 0.8), rel=1e-13)              # single panel (T <= 1)


@pytest.mark.parametrize("name", ["katb_brem", "brem_ps", "brem_vec", "brem_notl"])
def test_brem(hostcheck, name):
    g = load_golden(name)
    ma, gg, ns = float(g["ma"]), float(g["g"]), int(g["n_samples"])
    boson = str(g["boson_type"]) if "boson_type" in g else "pseudoscalar"
    tl = bool(g["use_track_length"]) if "use_track_length" in g else True
    C_ = O.Material("C")
    ef, pfx = g["electron_flux"], g["positron_flux"]
    par = P.brem_params(ma, gg, C_.z[0], C_.rad_length, ns, boson, 5.0, tl)
    rows = g["rows"]
    e0 = np.ascontiguousarray(ef[rows, 0]); w0 = np.ascontiguousarray(ef[rows, 1])
    u_rows = np.ascontiguousarray(np.nan_to_num(g["row_u"]))
    emit_ref = g["row_emit"].astype(bool)
    # the flat uniform log interleaves [row scalar][n_samples if emitted] per row: split it
    u = g["uniforms"]; pos = 0; samp = []
    for k in range(rows.shape[0]):
        if tl:
            assert u[pos] == g["row_u"][k]; pos += 1
        if emit_ref[k]:
            samp.append(u[pos:pos + ns]); pos += ns
    assert pos == u.shape[0]
    u_samp = np.ascontiguousarray(np.concatenate(samp)) if samp else np.zeros(0)
    n = int(emit_ref.sum()) * ns
    gx, gw = P.gauss_legendre()
    table = np.empty(rows.shape[0] * N.BREM_NP); emit = np.zeros(rows.shape[0], np.uint8)
    oe, of = np.empty(n), np.empty(n)
    ee, efl = np.ascontiguousarray(ef[:, 0]), np.ascontiguousarray(ef[:, 1])
    pe, pfl = np.ascontiguousarray(pfx[:, 0]), np.ascontiguousarray(pfx[:, 1])
    hostcheck.hc_brem(dptr(e0), dptr(w0), C.c_uint64(rows.shape[0]), dptr(u_rows), dptr(ee), dptr(efl), ee.shape[0],
                      dptr(pe), dptr(pfl), pe.shape[0], dptr(gx), dptr(gw), C.byref(par), C.c_uint32(ns), dptr(u_samp),
                      dptr(table), dptr(emit), dptr(oe), dptr(of))
    assert np.array_equal(emit.astype(bool), emit_ref)
    t = table.reshape(-1, N.BREM_NP)
    assert max_rel(t[:, 0], g["row_E1"]) == 0.0                       # E1 = ep_min + (E0-ep_min)*u bit exact
    assert max_rel(t[:, 1], g["row_W"]) < 1e-12                       # Gauss-Legendre vs QUADPACK
    assert max_rel(oe, g["axion_energy"]) < 1e-13
    assert max_rel(of, g["axion_flux"]) < TOL


def test_resonance(hostcheck):
    g = load_golden("katr_resonance")
    hostcheck.hc_resonance_sum.restype = C.c_double
    pf = g["positron_flux"]
    pe, pfl = np.ascontiguousarray(pf[:, 0]), np.ascontiguousarray(pf[:, 1])
    ma, ge, ns = 10.0, 1e-6, 1000
    eres = -O.M_E + ma ** 2 / (2 * O.M_E)
    emax = max(pf[:, 0])
    u = np.ascontiguousarray(g["uniforms"])
    s = hostcheck.hc_resonance_sum(dptr(pe), dptr(pfl), pe.shape[0], C.c_double(eres), C.c_double(emax), C.c_double(5.0),
                                   C.c_uint64(ns), dptr(u))
    C_ = O.Material("C")
    ntad = C_.rad_length * O.AVOGADRO / (C_.z[0] + C_.n[0])
    att = (5.0 * (emax - eres)) * s / ns
    wgt = C_.z[0] * (ntad * O.HBARC ** 2) * O.resonance_peak(ge, ma) * att
    assert wgt == pytest.approx(float(g["axion_flux"][0]), rel=1e-12)
    assert wgt == pytest.approx(185.58954304351653, rel=1e-12)


@pytest.mark.parametrize("name", ["kata_pairann", "pairann_mid"])
def test_pairann(hostcheck, name):
    g = load_golden(name)
    ma, ge, ns = float(g["ma"]), float(g["g"]), int(g["n_samples"])
    pf = g["positron_flux"]
    keep = O.pairann_row_valid(pf, ma)
    ep, wp = np.ascontiguousarray(pf[keep, 0]), np.ascontiguousarray(pf[keep, 1])
    C_ = O.Material("C")
    par = P.pairann_params(ma, ge, C_.z[0], C_.rad_length, ns)
    n = ep.shape[0] * ns
    u = np.ascontiguousarray(g["uniforms"])
    assert u.shape[0] == 2 * n
    oe, of = np.empty(n), np.empty(n)
    hostcheck.hc_pairann(dptr(ep), dptr(wp), C.c_uint64(ep.shape[0]), C.byref(par), C.c_uint32(ns), dptr(u), dptr(oe), dptr(of))
    assert max_rel(oe, g["axion_energy"]) < 1e-13
    assert max_rel(of, g["axion_flux"]) < TOL


def test_exp_fast_accuracy(hostcheck):
    """The fast-mode exp (Cody-Waite + degree-13 polynomial) against glibc: <= 1 ulp over the whole range used."""
    rs = np.random.RandomState(0)
    x = np.concatenate([-10 ** rs.uniform(-18, np.log10(708), 400000), -rs.uniform(0, 708, 400000),
                        rs.uniform(0, 708, 50000), [0.0, -0.0, -708.0, 708.0, -1e-300, -0.34657359, 0.34657359]])
    out = np.empty_like(x)
    hostcheck.hc_exp_fast(dptr(np.ascontiguousarray(x)), C.c_uint64(x.shape[0]), dptr(out))
    ref = np.exp(x)
    ulp = np.abs(out - ref) / np.spacing(ref)
    assert ulp.max() <= 1.0, (ulp.max(), x[np.argmax(ulp)])
    assert (ulp > 0).mean() < 0.25
    # outside [-708, 708] and NaN: library exp
    y = np.array([-709.5, -745.0, -800.0, 709.0, np.nan, -np.inf])
    o = np.empty_like(y)
    hostcheck.hc_exp_fast(dptr(y), C.c_uint64(y.shape[0]), dptr(o))
    with np.errstate(all="ignore"):
        e = np.exp(y)
    assert np.array_equal(o[~np.isnan(e)], e[~np.isnan(e)]) and np.isnan(o[4])


@pytest.mark.parametrize("fast", [0, 1])
def test_pairgamma(hostcheck, fast):
    g = load_golden("pairgamma")
    pf, ma, gg, ns = g["positron_flux"], float(g["ma"]), float(g["g"]), int(g["n_samples"])
    widths, centers = pf[1:, 0] - pf[:-1, 0], (pf[1:, 0] + pf[:-1, 0]) / 2
    ep_min = max((ma ** 2 - O.M_E ** 2) / (2 * O.M_E), O.M_E)
    keep = ~(centers < ep_min)
    par = N.PairGammaT()
    par.ma, par.ma2, par.ma4, par.me2, par.me4 = ma, ma ** 2, ma ** 4, O.M_E ** 2, O.M_E ** 4
    par.ep_min = ep_min
    par.nt = float(g["rad_length"]) * O.AVOGADRO / (2 * 6) * O.HBARC ** 2
    par.zc = np.int64(6) * 4 * np.pi * O.ALPHA * gg ** 2
    par.n_samples, par.max_t = float(ns), 5.0
    par.fast = fast
    nb = int(keep.sum()); n = nb * ns
    u = np.ascontiguousarray(g["uniforms"])
    assert u.shape[0] == 3 * n and 0 < nb <= centers.shape[0]
    oe, of = np.empty(n), np.empty(n)
    c, w = np.ascontiguousarray(centers[keep]), np.ascontiguousarray(widths[keep])
    pe, pfl = np.ascontiguousarray(pf[:, 0]), np.ascontiguousarray(pf[:, 1])
    hostcheck.hc_pairgamma(dptr(c), dptr(w), C.c_uint64(nb), C.c_uint32(ns), dptr(pe), dptr(pfl), pe.shape[0], C.byref(par),
                           dptr(np.ascontiguousarray(u[:2 * n])), dptr(np.ascontiguousarray(u[2 * n:])), dptr(oe), dptr(of))
    assert max_rel(oe, g["axion_energy"]) < 1e-13
    assert max_rel(of, g["axion_flux"]) < TOL


def test_general_decay(hostcheck):
    g = load_golden("decay4_angled")
    ma, ns = float(g["ma"]), int(g["n_samples"])
    ea, th, ph = g["axion_energy"], g["axion_angle"], g["parent_phis"]
    pa = np.sqrt(ea ** 2 - ma ** 2)
    p4 = np.ascontiguousarray(np.stack([ea, pa * np.cos(ph) * np.sin(th), pa * np.sin(ph) * np.sin(th), pa * np.cos(th)], axis=1))
    n = 3 * ns
    p1, p2, t1 = np.empty((4, n)), np.empty((4, n)), np.empty(n)
    hostcheck.hc_decay_general(dptr(p4), C.c_uint64(3), C.c_double(0.0), C.c_double(0.0), C.c_uint32(ns),
                               dptr(np.ascontiguousarray(g["uniforms"])), dptr(p1), dptr(p2), dptr(t1))
    scale = np.repeat(ea, ns)[None, :]
    assert np.max(np.abs(p1 - g["p1"].T) / scale) < 1e-13 and np.max(np.abs(p2 - g["p2"].T) / scale) < 1e-13
    m = load_golden("meson2body")
    mes = m["meson_flux"]
    p4 = np.ascontiguousarray(mes[:, [3, 0, 1, 2]])
    nm, ns = mes.shape[0], int(m["n_samples"])
    n = nm * ns
    p1, p2, t1 = np.empty((4, n)), np.empty((4, n)), np.empty(n)
    hostcheck.hc_decay_general(dptr(p4), C.c_uint64(nm), C.c_double(float(m["ma"])), C.c_double(0.0), C.c_uint32(ns),
                               dptr(np.ascontiguousarray(m["uniforms"])), dptr(p1), dptr(p2), dptr(t1))
    assert max_rel(p1[0], m["axion_energy"]) < 1e-12
    assert max_rel(t1, m["axion_angle"]) < TOL


# ---- charged-meson 3-body decays (alp_math3.cuh): the kernels' arithmetic, host build, against the reference fixtures ----
MESON3_MODELS = ["scalar_ib1", "pseudoscalar_ib1", "vector_ib1", "scalar_ib2", "vector_ib2", "vector_ib9", "sd", "vector_contact"]


def _meson3_par(mt, ma, model, c0, g, abd, ns):
    from alplib_b200 import _native as NV
    from oracle import alp_oracle as O
    mm, ckm, fM, tw = O.MESON_PARAMS[mt]
    p = NV.Meson3T()
    p.mm, p.ma, p.fM, p.ckm, p.total_width, p.c0, p.coupling = mm, ma, fM, ckm, tw, c0, g
    p.a, p.b, p.d = abd
    p.pref = (2 * mm) / (32 * np.power(2 * np.pi * mm, 3))
    p.n_samples = float(ns)
    p.model = MESON3_MODELS.index(model)
    return p, mm


def _meson3_rows(n_mesons, leptons):
    return (np.repeat(np.arange(n_mesons, dtype=np.uint32), len(leptons)),
            np.tile(np.asarray(leptons, dtype=np.float64), n_mesons))


def test_meson3_isotropic_host(hostcheck):
    from oracle import alp_oracle as O
    g = load_golden("meson3_iso")
    ns, c0, abd = int(g["n_samples"]), float(g["c0"]), tuple(float(x) for x in g["abd"])
    mf = np.ascontiguousarray(g["meson_flux"])
    worst = 0.0
    for k in range(int(g["n_cases"])):
        mi, mt, ma, cpl = g[f"c{k}_cfg"]
        par, mm = _meson3_par(["pion", "kaon"][int(mt)], float(ma), MESON3_MODELS[int(mi)], c0, float(cpl), abd, ns)
        leptons = [ml for ml in (O.M_E, O.M_MU) if not (ma > mm - ml)]
        rm, rl = _meson3_rows(mf.shape[0], leptons)
        n = rm.shape[0] * ns
        oe, of = np.empty(n), np.empty(n)
        u = np.ascontiguousarray(g[f"c{k}_u"])
        hostcheck.hc_meson3_isotropic(dptr(mf), dptr(rm), dptr(rl), C.c_uint64(rm.shape[0]), C.c_uint32(ns), C.byref(par),
                                      dptr(u), dptr(oe), dptr(of))
        assert max_rel(oe, g[f"c{k}_E"]) < 1e-14
        worst = max(worst, max_rel(of, g[f"c{k}_F"]))
    assert worst < 1e-9, worst


def test_meson3_decay_host(hostcheck):
    from oracle import alp_oracle as O
    g = load_golden("meson3_decay")
    ns = int(g["n_samples"])
    mf = np.ascontiguousarray(g["meson_flux"])
    for k in range(int(g["n_cases"])):
        mi, mt, ma, cpl, cut = g[f"c{k}_cfg"]
        par, mm = _meson3_par(["pion", "kaon"][int(mt)], float(ma), MESON3_MODELS[int(mi)], float(g["c0"]), float(cpl), (0, 0, 0), ns)
        leptons = [ml for ml in (O.M_E, O.M_MU) if not (ma > mm - ml)]
        rm, rl = _meson3_rows(mf.shape[0], leptons)
        n = rm.shape[0] * ns
        outs = [np.empty(n) for _ in range(6)]
        keep = np.zeros(n, dtype=np.uint8)
        up, us = np.ascontiguousarray(g[f"c{k}_upos"]), np.ascontiguousarray(g[f"c{k}_usamp"])
        hostcheck.hc_meson3_decay(dptr(mf), dptr(rm), dptr(rl), C.c_uint64(rm.shape[0]), C.c_uint32(ns), C.byref(par),
                                  C.c_double(float(g["det_dist"])), C.c_double(float(g["det_length"])),
                                  C.c_double(float(g["energy_cut"])), C.c_int(int(cut)), dptr(up), dptr(us),
                                  *[dptr(o) for o in outs], dptr(keep))
        sel = keep != 0
        for o, key, tol in zip(outs, ("energy", "flux", "cosine", "angle", "solid_angle", "decay_pos"),
                               (1e-13, 1e-9, 1e-12, 1e-9, 1e-14, 1e-14)):
            assert max_rel(o[sel], g[f"c{k}_{key}"]) < tol, (k, key)


def test_meson3_quadrature_refines_like_quad(hostcheck):
    """Couplings of order one push the Dalitz integral above quad's epsabs: the kernels' QAGS-style bisection must follow."""
    from oracle import alp_oracle as O
    rng = np.random.default_rng(4)
    for model, mt, ma, ml in (("vector_ib1", "kaon", 0.1, O.M_E), ("vector_ib9", "kaon", 5.0, O.M_MU), ("sd", "kaon", 30.0, O.M_E)):
        par, mm = _meson3_par(mt, ma, model, -0.95, 1.0, (1.0, 0.5, 0.2), 10)
        ea = ma + ((mm ** 2 + ma ** 2 - ml ** 2) / (2 * mm) - ma) * rng.random(60)
        out = np.empty_like(ea)
        hostcheck.hc_meson3_dgamma(dptr(ea), C.c_uint64(ea.size), C.c_double(ml), C.byref(par), dptr(out))
        _, ckm, fM, _ = O.MESON_PARAMS[mt]
        ref = np.array([O.meson3_dgamma_dea(e, ml, ma, mm, fM, ckm, model, -0.95, 1.0, (1.0, 0.5, 0.2)) for e in ea])
        assert max_rel(out, ref) < 1e-9


def test_detection_cross_sections_fast_mode(hostcheck):
    """Fast-mode inverse-Compton / inverse-Primakoff cross sections against the oracle over the whole energy range, including
    the samples that fall back to the reference expression (E_a <= m_a, below threshold, nan_to_num corners)."""
    rng = np.random.default_rng(8)
    ev = P.event_params(1.0 / 86400.0, 0.0)
    for ma in (0.01, 0.5, 1.5, 30.0):
        e = np.concatenate([10 ** rng.uniform(np.log10(ma) - 0.5, 6, 4000), [ma, ma * (1 + 1e-12), 0.999 * ma, 1e7]])
        w = np.ones(e.shape[0], dtype=np.float32)
        out = np.empty(e.shape[0])
        xs = P.icompton_consts(ma, 1e-5, np.int64(18))
        hostcheck.hc_scatter(2, dptr(e), dptr(w), C.c_uint64(e.shape[0]), xs, C.c_double(1.0), C.c_double(1.0), C.byref(ev), dptr(out))
        with np.errstate(all="ignore"):
            ref = O.icompton_sigma(e, ma, 1e-5, 18)
        assert max_rel(out, ref, floor=1e-300) < TOL, ma
        xs = P.iprimakoff_consts(ma, 1e-5, np.int64(18))
        hostcheck.hc_scatter(3, dptr(e), dptr(w), C.c_uint64(e.shape[0]), xs, C.c_double(1.0), C.c_double(1.0), C.byref(ev), dptr(out))
        with np.errstate(all="ignore"):
            ref = O.iprimakoff_sigma(e, 1e-5, ma, 18)
        assert max_rel(out, ref, floor=1e-300) < TOL, ma


def test_brem_fast_production(hostcheck):
    """Fast-mode bremsstrahlung sample (table 10^x, shared reciprocals) against the reference-order expression on random rows,
    pseudoscalar and vector, including the end-point region that falls back to the reference expression."""
    rng = np.random.default_rng(12)
    hostcheck.hc_exp10_fast.restype = C.c_double
    x = rng.uniform(-300, 300, 20000)
    y = np.array([hostcheck.hc_exp10_fast(C.c_double(v)) for v in x[:4000]])
    assert max_rel(y, np.power(10.0, x[:4000])) < 4e-16
    n = 40000
    for vector in (0, 1):
        for ma in (0.05, 10.0, 300.0):
            from alplib_b200.materials import Material
            mat = Material("C")
            par = P.brem_params(ma, 1e-6, mat.z[0], mat.rad_length, 100, "vector" if vector else "pseudoscalar", 5.0, True)
            e1 = 10 ** rng.uniform(np.log10(max(ma, 0.511)) + 0.01, 5.5, n)
            ea_max = e1 * (1 - (ma / e1) ** 2)
            ok = ea_max > ma
            e1, ea_max = e1[ok], ea_max[ok]
            m = e1.shape[0]
            rows = np.zeros((m, N.BREM_NP))
            rows[:, 0] = e1; rows[:, 1] = 10 ** rng.uniform(5, 12, m)
            rows[:, 2] = np.log10(ea_max) - np.log10(ma)
            rows[:, 3] = par.pref_num / e1
            rows[:, 4] = 1.0
            rows[:, 5] = np.log10(ea_max / e1) - np.log10(ma / e1)
            u = np.concatenate([rng.random(m - 200), 1 - 10 ** rng.uniform(-12, -3, 200)])      # some at the end point
            ref, fast = np.empty(2 * m), np.empty(2 * m)
            hostcheck.hc_brem_sample_pair(dptr(np.ascontiguousarray(rows)), dptr(u), C.c_uint64(m), C.byref(par), dptr(ref), dptr(fast))
            assert max_rel(fast[:m], ref[:m]) < 1e-14, (vector, ma)
            assert max_rel(fast[m:], ref[m:], floor=1e-300) < TOL, (vector, ma)


def test_pairann_fast_production(hostcheck):
    """Fast-mode associated-production sample (cos theta = 1 - 2u, one shared reciprocal) against the reference-order one."""
    from alplib_b200.materials import Material
    rng = np.random.default_rng(21)
    mat = Material("C")
    for ma in (0.05, 10.0, 60.0):
        par = P.pairann_params(ma, 1e-6, mat.z[0], mat.rad_length, 100)
        thr = max((ma ** 2 - 2 * 0.511 ** 2) / (2 * 0.511), 0.511)
        n = 30000
        ep = thr * (1 + 10 ** rng.uniform(-6, 4, n))
        wgt = 10 ** rng.uniform(5, 12, n)
        u = rng.random(n)
        ref, fast = np.empty(2 * n), np.empty(2 * n)
        hostcheck.hc_pairann_sample_pair(dptr(ep), dptr(wgt), dptr(u), C.c_uint64(n), C.byref(par), dptr(ref), dptr(fast))
        ok = np.isfinite(ref[n:])
        assert ok.mean() > 0.9
        assert max_rel(fast[:n][ok], ref[:n][ok]) < 1e-12, ma
        assert max_rel(fast[n:][ok], ref[n:][ok], floor=1e-300) < TOL, ma


def test_resonance_fast_term(hostcheck):
    """Fast-mode resonance term (exp-table pow, polynomial 1/Gamma, spacing-guessed np.interp) against the reference-order
    term, uniform and non-uniform positron tables; 1/Gamma against scipy over the range reached (and beyond)."""
    from scipy.special import gamma
    hostcheck.hc_rgamma_fast.restype = C.c_double
    z = np.concatenate([10 ** np.random.default_rng(2).uniform(-8, np.log10(30.0), 4000), [1.0, 2.0, 3.0, 1.5, 6.6667, 1e-12]])
    r = np.array([hostcheck.hc_rgamma_fast(C.c_double(v)) for v in z])
    assert max_rel(r, 1.0 / gamma(z)) < 1e-14
    rng = np.random.default_rng(5)
    n = 50000
    u = rng.random(2 * n)
    for uniform in (True, False):
        pe = np.linspace(10.0, 5000.0, 200) if uniform else np.sort(10 ** rng.uniform(1, 3.7, 200))
        pfl = 3e11 * np.exp(-pe / 600.0)
        inv_dx = 1.0 / ((pe[-1] - pe[0]) / (pe.shape[0] - 1)) if uniform else 0.0
        for eres, emax, max_t in ((97.3, 5000.0, 5.0), (10.0, 4000.0, 5.0), (2000.0, 6000.0, 3.0)):
            ref, fast = np.empty(n), np.empty(n)
            hostcheck.hc_resonance_terms(dptr(pe), dptr(pfl), pe.shape[0], C.c_double(eres), C.c_double(emax), C.c_double(max_t),
                                         C.c_uint64(n), dptr(u), C.c_double(inv_dx), dptr(ref), dptr(fast))
            assert max_rel(fast, ref, floor=1e-300) < TOL, (uniform, eres)
            assert abs(fast.sum() / ref.sum() - 1) < 1e-13


@pytest.mark.parametrize("regime", ["long_lived", "short_lived", "stable", "very_short"])
def test_propagate_fast_regimes(hostcheck, regime):
    """The fast propagate in every lifetime regime against the oracle: long-lived (single range test), short-lived (each
    exponential on its own, exact zeros below -746, library path only in the denormal band), stable (width 0) and so
    short-lived that every sample is fully decayed -- energies from threshold to 1e6 MeV, element by element."""
    ma = 10.0
    width = {"long_lived": 1e-22, "short_lived": 4e-13, "stable": 0.0, "very_short": 1e-3}[regime]
    rng = np.random.default_rng(17)
    e = np.concatenate([ma * (1 + 10 ** rng.uniform(-9, 0, 3000)), 10 ** rng.uniform(np.log10(2 * ma), 6, 3000)])
    w = 10 ** rng.uniform(-5, 8, e.shape[0])
    prop = P.prop_params(ma, width, 1.7, 574.0, 5.0, geom_accept=GA, fast=True)
    d32, s32, d64, s64 = run_propagate(hostcheck, e, w, prop)
    with np.errstate(all="ignore"):
        p = O.propagate(e, w, ma, width, 1.7, 574.0, 5.0)
    assert max_rel(s64, p["scatter_w64"], floor=1e-300) < TOL
    assert_decay_close(d64, p["decay_w64"], p["scatter_w64"], TOL)
    if regime == "very_short":
        assert np.all(s64 == 0) and np.all(p["scatter_w64"] == 0)
    if regime == "short_lived":
        assert 0 < np.count_nonzero(s64 == 0) < e.shape[0]          # both the exact-zero and the regular branch are exercised


def test_seed21_flavour_runs_the_device_corrections():
    """The two host builds differ only in how the fast mode's reciprocal / reciprocal square root are formed: "seed21" must
    really go through the Newton steps on a truncated seed (results differ from the exact ones) and stay far inside 1e-9."""
    import ctypes
    import os
    import subprocess
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck")
    libs = {}
    for name, flags in (("libhostcheck.so", []), ("libhostcheck_seed21.so", ["-DALP_EMULATE_APPROX_SEED"])):
        so = os.path.join(d, name)
        if not os.path.exists(so):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared"] + flags
                                  + ["-o", so, os.path.join(d, "hostcheck.cpp")])
        libs[name] = ctypes.CDLL(so)
    rng = np.random.default_rng(5)
    e = 1.5 * (1 + 10 ** rng.uniform(-6, 3, 4000))
    w = 10 ** rng.uniform(-5, 8, e.shape[0])
    prop = P.prop_params(1.5, 1e-18, 1.0, 4.0, 0.2, geom_accept=GA, fast=True)
    out = {k: run_propagate(lib, e, w, prop) for k, lib in libs.items()}
    a, b = out["libhostcheck.so"][3], out["libhostcheck_seed21.so"][3]          # float64 scatter weights
    rel = np.abs(a - b) / np.abs(a)
    assert 1e-17 < rel.max() < 1e-11, rel.max()

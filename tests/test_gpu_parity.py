"""GPU tier (pytest -m gpu, runs on the B200 box): the CUDA path, called through the reference-shaped Python classes
and the C ABI of libalpk.so, against (i) the golden fixtures produced by the unmodified reference and (ii) the NumPy
oracle on the same injected uniforms.  Tolerance: relative 1e-9 on FP64 quantities (north star), 1.2e-7 (one float32
ulp) on the reference's float32 attributes; decay weights use the conditioned tolerance of conftest.decay_tol."""
import numpy as np
import pytest

from conftest import assert_decay_close, decay_tol, load_golden, max_rel
from oracle import alp_oracle as O
from oracle import philox as PX

pytestmark = pytest.mark.gpu

TOL = 1e-9
F32 = 1.2e-7
GEOM = dict(det_dist=4.0, det_length=0.2, det_area=0.04)
GA = O.geom_acceptance(0.04, 4.0)


@pytest.fixture(autouse=True, params=["fast", "faithful"])
def precision(request):
    """Every GPU parity test runs in both arithmetic modes (alplib_b200.config.PRECISION)."""
    from alplib_b200 import config
    old, old_rounds = config.PRECISION, PX.PRODUCTION_ROUNDS
    config.PRECISION = request.param
    PX.PRODUCTION_ROUNDS = config.PHILOX_ROUNDS[request.param]      # the oracle's Philox replays the stream of the mode under test
    yield request.param
    config.PRECISION, PX.PRODUCTION_ROUNDS = old, old_rounds


@pytest.fixture(scope="module")
def A():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    import alplib_b200
    from alplib_b200 import _native
    _native.lib()          # fails loudly if libalpk.so is missing
    return alplib_b200


@pytest.mark.parametrize("mat", ["W", "C"])
def test_abs_sigma(A, mat):
    g = load_golden("abs_table_" + mat)
    xs = A.AbsCrossSection(A.Material(mat))
    assert max_rel(xs.sigma_mev(g["energy"]), g["sigma_mev"]) < 1e-12
    assert float(xs.sigma_mev(2e5)) == pytest.approx(2.5680599121454694e+21, rel=1e-12)   # out-of-range quirk kept


def test_cross_section_class_methods(A):
    """sigma_cm2 / sigma_mev / mu of AbsCrossSection and ALPPairProdutionCrossSection against the reference (photon_xs.py)."""
    g = load_golden("xs_methods")
    e = g["energy"]
    for mat in ("W", "Ar", "NaI"):
        a = A.AbsCrossSection(A.Material(mat))
        assert max_rel(a.sigma_cm2(e), g[f"abs_{mat}_cm2"]) < 1e-12 and max_rel(a.sigma_mev(e), g[f"abs_{mat}_mev"]) < 1e-12
        assert max_rel(a.mu(e, 6.3e22), g[f"abs_{mat}_mu"]) < 1e-12
        assert float(a.sigma_cm2(5.0)) == pytest.approx(float(g[f"abs_{mat}_cm2"][5]), rel=1e-12)
    for tag, mat, ma in (("ar01", "Ar", 0.1), ("ar095", "Ar", 0.95), ("ge2", "Ge", 2.0)):
        x = A.ALPPairProdutionCrossSection(A.Material(mat), ma=ma)
        assert max_rel(x.sigma_cm2(e), g[f"pair_{tag}_cm2"]) < 1e-12 and max_rel(x.sigma_mev(e), g[f"pair_{tag}_mev"]) < 1e-12
        assert max_rel(x.mu(e, 2.1e22), g[f"pair_{tag}_mu"]) < 1e-12
        assert x.sigma_mev(0.9) == 0.0                        # below 2 m_e


def test_c1_readme_example(A):
    """BASELINE config 1: README Primakoff example end to end."""
    g = load_golden("c1_primakoff")
    fl = A.FluxPrimakoffIsotropic(g["photon_flux"], A.Material("W"), axion_mass=0.1, axion_coupling=1e-5,
                                  n_samples=1000, keep_f64=True, **GEOM)
    fl.simulate()
    assert list(fl.axion_energy) == [100.0]
    assert max_rel(fl.axion_flux, [327952.3965351194]) < TOL
    fl.propagate()
    assert fl.decay_axion_weight.dtype == np.float32 and fl.scatter_axion_weight.dtype == np.float32
    assert max_rel(fl.decay_axion_weight, g["decay_w32"]) <= F32
    assert max_rel(fl.scatter_axion_weight, g["scatter_w32"]) <= F32
    assert max_rel(fl.decay_axion_weight_f64 * GA, g["decay_w64"]) < TOL
    gen = A.PhotonEventGenerator(fl, A.Material("Ar"))
    rate = gen.decays(100, 5.0)
    assert isinstance(rate, np.float64)
    assert rate == pytest.approx(284.1535949707031, rel=2 * F32)
    assert max_rel(gen.decay_weights, g["decay_weights"]) <= 2 * F32
    pending = gen.decays_async(100, 5.0)                    # the same rate, read late (pinned copy + event)
    later = A.PhotonEventGenerator(fl, A.Material("Ar")).decays(50, 5.0)          # other work queued in between
    assert pending.result() == rate and float(pending) == float(rate) and pending.done() and later < rate
    ipr = gen.inverse_primakoff(1e-5, 0.1, 1e27, 100, 5.0)
    assert ipr == pytest.approx(float(g["iprimakoff_1e27_100_5"]), rel=2 * F32)
    # 1-D photon table fails like the reference (README.md:90 quirk)
    with pytest.raises(IndexError):
        A.FluxPrimakoffIsotropic(np.array([100.0, 1e12]), A.Material("W")).simulate()
    with pytest.raises(Exception, match="No such detector"):
        A.Material("Unobtainium")


def test_primakoff_ragged(A):
    g = load_golden("primakoff_ragged")
    ma, gg = float(g["ma"]), float(g["g"])
    fl = A.FluxPrimakoffIsotropic(g["photon_flux"], A.Material("W"), axion_mass=ma, axion_coupling=gg, keep_f64=True, **GEOM)
    fl.simulate()
    assert np.array_equal(fl.axion_energy, g["axion_energy"])           # rows below ma dropped, order kept
    assert max_rel(fl.axion_flux, g["axion_flux"]) < TOL
    assert len(fl.decay_axion_weight) == 0                                # reset by simulate()
    fl.propagate()
    assert max_rel(fl.scatter_axion_weight_f64 * GA, g["scatter_w64"]) < TOL
    assert_decay_close(fl.decay_axion_weight_f64 * GA, g["decay_w64"], g["scatter_w64"], TOL)
    assert max_rel(fl.scatter_axion_weight, g["scatter_w32"]) <= F32
    assert_decay_close(fl.decay_axion_weight, g["decay_w32"], g["scatter_w32"], TOL, extra_rel=F32)
    fl.propagate(new_coupling=float(g["new_coupling"]))
    assert max_rel(fl.scatter_axion_weight_f64 * GA, g["scatter_w64_new"]) < TOL
    assert_decay_close(fl.decay_axion_weight_f64 * GA, g["decay_w64_new"], g["scatter_w64_new"], TOL)
    rate = A.PhotonEventGenerator(fl, A.Material("Ar")).decays(30, 1.0)
    assert rate == pytest.approx(float(g["decays_30_1"]), rel=1e-6)


@pytest.mark.parametrize("name", ["katc_compton", "compton_mid", "compton_loop"])
def test_compton_chain_injected(A, name):
    g = load_golden(name)
    ma, gg, ns = float(g["ma"]), float(g["g"]), int(g["n_samples"])
    loop = bool(g["loop_decay"]) if "loop_decay" in g else False
    iso = bool(g["is_isotropic"]) if "is_isotropic" in g else True
    fl = A.FluxComptonIsotropic(g["photon_flux"], A.Material("W"), axion_mass=ma, axion_coupling=gg, n_samples=ns,
                                loop_decay=loop, is_isotropic=iso, keep_f64=True, **GEOM)
    fl.simulate(uniforms=g["uniforms"])
    assert np.array_equal(fl.axion_energy, g["axion_energy"])             # ma + (eg-ma)*u: bit exact
    assert max_rel(fl.axion_flux, g["axion_flux"]) < TOL
    fl.propagate()
    ga = GA if iso else 1.0
    assert max_rel(fl.scatter_axion_weight_f64 * ga, g["scatter_w64"]) < TOL
    assert_decay_close(fl.decay_axion_weight_f64 * ga, g["decay_w64"], g["scatter_w64"], TOL)
    assert max_rel(fl.scatter_axion_weight, g["scatter_w32"]) <= F32
    assert_decay_close(fl.decay_axion_weight, g["decay_w32"], g["scatter_w32"], TOL, extra_rel=F32)
    gen = A.ElectronEventGenerator(fl, A.Material("Ar"))
    r_dec = gen.decays(100, 5.0)
    r_com = gen.compton(gg, ma, 1e27, 100, 5.0)
    assert r_com == pytest.approx(float(g["compton_1e27_100_5"]), rel=1e-6)
    if float(g["decays_100_5"]) != 0:
        assert r_dec == pytest.approx(float(g["decays_100_5"]), rel=1e-4 if name == "compton_loop" else 1e-6)
    if "scatter_weights" in g:
        assert max_rel(gen.scatter_weights, g["scatter_weights"], floor=1e-300) < 2 * F32
    if "pair_weights" in g:
        r_pair = gen.pair_production(gg, 1e27, 100, 1.0)
        assert r_pair == pytest.approx(float(g["pair_1e27_100_1"]), rel=1e-6) and r_pair > 0
        assert max_rel(gen.pair_weights, g["pair_weights"], floor=1e-300) < 2 * F32
    if "new_coupling" in g:
        fl.propagate(new_coupling=float(g["new_coupling"]))
        assert_decay_close(fl.decay_axion_weight_f64 * GA, g["decay_w64_new"], g["scatter_w64_new"], TOL)
        assert_decay_close(fl.decay_axion_weight, g["decay_w32_new"], g["scatter_w32_new"], TOL, extra_rel=F32)


def _c2_flux(n_bins):
    e = np.logspace(np.log10(0.5), np.log10(1e3), n_bins)
    return np.stack([e, 1e16 * e ** -1.5 * np.exp(-e / 150)], axis=1)


@pytest.mark.parametrize("ns", [501, 1000, 20000])      # odd: pairs straddle rows; even: paired kernels, short / long rows
def test_compton_free_running_matches_oracle_sample_by_sample(A, ns):
    """Free-running Philox mode: replay the kernel's uniforms with the NumPy Philox and feed them to the oracle."""
    pf = _c2_flux(37)
    ma, gg, seed = 1.5, 1e-7, 0xA1B2C3D4 + 2
    fl = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=ma, axion_coupling=gg, n_samples=ns, seed=seed,
                                keep_f64=True, **GEOM)
    fl.simulate()
    fl.propagate()
    keep = O.compton_row_valid(pf, ma)
    assert 0 < keep.sum() < pf.shape[0]
    u = np.concatenate([PX.uniform_1(seed, PX.TAG_COMPTON_EA, int(r), ns) for r in np.nonzero(keep)[0]])
    oE, oF = O.compton_simulate(pf, ma, gg, 74, O.AbsCrossSection(O.Material("W")), ns, O.UniformSource(u))
    assert np.array_equal(fl.axion_energy, oE)
    assert max_rel(fl.axion_flux, oF) < TOL
    p = O.propagate(oE, oF, ma, O.W_ee(gg, ma), 1.0, 4.0, 0.2, geom_accept=GA)
    assert max_rel(fl.scatter_axion_weight_f64, p["scatter_w64"]) < TOL
    assert_decay_close(fl.decay_axion_weight_f64, p["decay_w64"], p["scatter_w64"], TOL)
    # float32 attribute = float32(w64) * float32(acceptance): two roundings, so FP64 values that agree to 1e-15 but straddle a
    # float32 rounding boundary can end up two float32 ulps apart (seen once in 5.4e5 samples)
    assert_decay_close(fl.decay_axion_weight, p["decay_w"], p["scatter_w"], TOL, extra_rel=2 * F32)
    _, orate = O.decays(oE, p["decay_w"], 100, 5.0)
    rate = A.ElectronEventGenerator(fl, A.Material("Ar")).decays(100, 5.0)
    assert rate == pytest.approx(orate, rel=1e-6)


def test_primakoff_fused_equals_unfused(A):
    """The one-kernel Primakoff schedule (alpk_primakoff_chain: production -> propagate -> decay weights + rate) leaves the
    arrays and the rate of the call-by-call schedule, bit for bit, in every arithmetic mode; fused=True materialises nothing
    until an attribute is read."""
    pf = _c2_flux(3000)
    for precision in ("fast", "faithful", "fp32"):
        res = {}
        for fused in (False, "auto", True):
            fl = A.FluxPrimakoffIsotropic(pf, A.Material("W"), axion_mass=0.1, axion_coupling=1e-5, fused=fused, keep_f64=(fused != True),
                                          precision=precision, **GEOM)
            fl.simulate(); fl.propagate(new_coupling=3e-5)
            gen = A.PhotonEventGenerator(fl, A.Material("Ar"))
            rate = gen.decays(100, 5.0)
            if fused is True:
                assert fl._E is None and fl._D32 is None            # reduce-only: nothing materialised yet
            res[fused] = (rate, np.asarray(fl.axion_energy), np.asarray(fl.axion_flux), np.asarray(fl.decay_axion_weight),
                          np.asarray(fl.scatter_axion_weight), np.asarray(gen.decay_weights))
            if fused == "auto":
                assert np.array_equal(fl.decay_axion_weight_f64, ref_f64)
            if fused is False:
                ref_f64 = np.asarray(fl.decay_axion_weight_f64)
        for fused in ("auto", True):
            assert res[fused][0] == res[False][0], (precision, fused)
            for a, b in zip(res[fused][1:], res[False][1:]):
                assert a.shape == b.shape and np.array_equal(a, b), (precision, fused)
        # explicit one-kernel call
        fl = A.FluxPrimakoffIsotropic(pf, A.Material("W"), axion_mass=0.1, axion_coupling=1e-5, precision=precision, **GEOM)
        fl.simulate()
        r, evw = fl.chain(100, 5.0, new_coupling=3e-5)
        assert float(r.item()) == res[False][0] and np.array_equal(evw.cpu().numpy(), res[False][5])
        r2, none = fl.chain(100, 5.0, new_coupling=3e-5, materialize=False)
        assert none is None and float(r2.item()) == res[False][0]


def test_compton_fused_equals_unfused(A):
    """One-kernel chain (materialising and reduce-only) vs the three separate launches: identical samples."""
    pf = _c2_flux(300)
    kw = dict(axion_mass=1.5, axion_coupling=1e-7, n_samples=1000, seed=77, **GEOM)
    a = A.FluxComptonIsotropic(pf, A.Material("W"), fused=False, **kw)
    a.simulate(); a.propagate(new_coupling=3e-7)
    gen_a = A.ElectronEventGenerator(a, A.Material("Ar"))
    ra = gen_a.decays(100, 5.0)
    b = A.FluxComptonIsotropic(pf, A.Material("W"), **kw)
    b.simulate()
    rate_b, evw_b = b.chain(100, 5.0, new_coupling=3e-7, materialize=True)
    assert np.array_equal(a.axion_energy, b.axion_energy) and np.array_equal(a.axion_flux, b.axion_flux)
    assert np.array_equal(a.decay_axion_weight, b.decay_axion_weight)
    assert np.array_equal(a.scatter_axion_weight, b.scatter_axion_weight)
    assert np.array_equal(gen_a.decay_weights, evw_b.cpu().numpy())
    assert float(rate_b.item()) == pytest.approx(ra, rel=1e-12)
    c = A.FluxComptonIsotropic(pf, A.Material("W"), fused=True, **kw)
    c.simulate(); c.propagate(new_coupling=3e-7)
    assert c._E is None and c._D32 is None                         # nothing launched yet
    rc = A.ElectronEventGenerator(c, A.Material("Ar")).decays(100, 5.0)
    assert c._E is None                                            # reduce-only: still nothing materialised
    assert rc == pytest.approx(ra, rel=1e-12)
    assert np.array_equal(c.axion_energy, a.axion_energy)          # recomputed on demand, bit-identical
    assert np.array_equal(c.decay_axion_weight, a.decay_axion_weight)
    # default schedule ("auto"): deferred, then ONE kernel at decays() that materialises every array
    d = A.FluxComptonIsotropic(pf, A.Material("W"), **kw)
    assert d.fused == "auto"
    d.simulate(); d.propagate(new_coupling=3e-7)
    assert d._E is None and d._D32 is None
    gen_d = A.ElectronEventGenerator(d, A.Material("Ar"))
    rd = gen_d.decays(100, 5.0)
    assert d._E is not None and d._D32 is not None and rd == pytest.approx(ra, rel=1e-12)
    assert np.array_equal(d.axion_energy, a.axion_energy) and np.array_equal(d.axion_flux, a.axion_flux)
    assert np.array_equal(d.decay_axion_weight, a.decay_axion_weight) and np.array_equal(d.scatter_axion_weight, a.scatter_axion_weight)
    assert np.array_equal(gen_d.decay_weights, gen_a.decay_weights)
    # reading an attribute before decays() launches just what it needs
    e_ = A.FluxComptonIsotropic(pf, A.Material("W"), **kw)
    e_.simulate()
    assert np.array_equal(e_.axion_flux, a.axion_flux) and len(e_.decay_axion_weight) == 0
    e_.propagate(new_coupling=3e-7)
    assert np.array_equal(e_.scatter_axion_weight, a.scatter_axion_weight)


@pytest.mark.parametrize("ns,n_bins", [(256, 40), (1000, 300), (10000, 37), (4098, 11)])
def test_row_kernel_equals_tile_kernel(A, ns, n_bins, monkeypatch):
    """The row-aligned warp-unit kernel (chain_rows.cuh, dynamic and static unit assignment) and the tile kernel
    (chain_kernel.cuh) give the same bits for every per-sample output, the same rate to summation order, and -- unit sums
    being added in unit order -- the row kernel's rate does not depend on how the units were scheduled."""
    pf = _c2_flux(n_bins)
    kw = dict(axion_mass=1.5, axion_coupling=1e-7, n_samples=ns, seed=4242, **GEOM)
    edges = np.logspace(np.log10(1.5), 3, 33)
    res = {}
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("ALPK_CHAIN_ROWS", mode)
        fl = A.FluxComptonIsotropic(pf, A.Material("W"), **kw)
        fl.row_offset = 17
        fl.simulate()
        rate, evw = fl.chain(100, 5.0, new_coupling=3e-7, materialize=True)
        rate_ro, _ = fl.chain(100, 5.0, new_coupling=3e-7, materialize=False)
        rate_h, _ = fl.chain(100, 5.0, new_coupling=3e-7, materialize=False, hist_edges=edges)
        hist = fl.last_hist.cpu().numpy()
        prod = A.FluxComptonIsotropic(pf, A.Material("W"), fused=False, **kw)      # production-only and production+propagate launches
        prod.row_offset = 17
        prod.simulate()
        res[mode] = (fl.axion_energy.copy(), fl.axion_flux.copy(), fl.decay_axion_weight.copy(), fl.scatter_axion_weight.copy(),
                     evw.cpu().numpy(), float(rate.item()), float(rate_ro.item()), prod.axion_flux.copy(), float(rate_h.item()), hist)
    ref = res["0"]
    assert ref[0].shape[0] > 0
    for mode in ("1", "2"):
        for k in (0, 1, 2, 3, 4, 7):
            assert np.array_equal(res[mode][k], ref[k]), (mode, k)
        assert res[mode][5] == pytest.approx(ref[5], rel=1e-12)
        assert np.allclose(res[mode][9], ref[9], rtol=1e-11, atol=0.0) and res[mode][9].sum() == pytest.approx(ref[5], rel=1e-9)
        # materialising and reduce-only launches sum the same unit sums in the same order (the histogram launch runs on the
        # tile kernel: same rate to summation order)
        assert res[mode][5] == res[mode][6] and res[mode][8] == pytest.approx(ref[5], rel=1e-12)
    assert res["1"][5] == res["2"][5]            # dynamic vs static unit assignment: identical rate, bit for bit


def test_strict_compat_containers_and_lazy_decay_weights(A):
    """strict=True materialises the reference's host containers -- Python lists of np.float64 after simulate() (fluxes.py:222-224)
    and three Python lists from simulate_decay_4vectors (generators.py:124-128) -- and in the reduce-only schedule
    decay_weights is computed on demand (it is what the reference's decays() leaves behind, generators.py:88-90)."""
    pf = _c2_flux(60)
    kw = dict(axion_mass=1.5, axion_coupling=1e-7, n_samples=64, seed=9, **GEOM)
    ref = A.FluxComptonIsotropic(pf, A.Material("W"), fused=False, **kw)
    ref.simulate(); ref.propagate()
    gen_ref = A.ElectronEventGenerator(ref, A.Material("Ar"))
    rate_ref = gen_ref.decays(100, 5.0)
    fl = A.FluxComptonIsotropic(pf, A.Material("W"), strict=True, **kw)
    fl.simulate()
    assert isinstance(fl.axion_energy, list) and isinstance(fl.axion_flux, list) and isinstance(fl.axion_energy[0], np.float64)
    assert fl.axion_energy == list(ref.axion_energy) and fl.axion_flux == list(ref.axion_flux)
    assert (fl.axion_energy + [1.0])[-1] == 1.0                     # list semantics: concatenation, not broadcasting
    fl.propagate()
    assert isinstance(fl.decay_axion_weight, np.ndarray) and fl.decay_axion_weight.dtype == np.float32   # as after the reference's propagate()
    # reduce-only schedule: decays() launches one kernel and keeps nothing; reading decay_weights afterwards computes them
    fr = A.FluxComptonIsotropic(pf, A.Material("W"), fused=True, **kw)
    fr.simulate(); fr.propagate()
    gen = A.ElectronEventGenerator(fr, A.Material("Ar"))
    assert gen.decays(100, 5.0) == pytest.approx(rate_ref, rel=1e-12) and fr._E is None
    assert np.array_equal(gen.decay_weights, gen_ref.decay_weights) and np.count_nonzero(gen.decay_weights) > 0
    r_dev = gen.decays_device(50, 5.0)                                # a later call must not leave the earlier weights behind
    assert np.allclose(gen.decay_weights, 0.5 * gen_ref.decay_weights, rtol=1e-6) and float(r_dev.item()) == pytest.approx(0.5 * rate_ref, rel=1e-6)
    fr.propagate(new_coupling=2e-7)
    assert np.allclose(gen.decay_weights, 0.5 * gen_ref.decay_weights, rtol=1e-6)     # already materialised: what that decays() left behind
    gen.decays_device(25, 5.0)
    fr.propagate(new_coupling=3e-7)
    with pytest.raises(RuntimeError):
        gen.decay_weights                                              # not materialised and the flux changed after decays()
    # decay MC: three Python lists in strict mode
    prim = A.FluxPrimakoffIsotropic(np.array([[100.0, 1e12], [40.0, 1e11]]), A.Material("W"), axion_mass=0.1, axion_coupling=1e-5,
                                    strict=True, **GEOM)
    prim.simulate(); prim.propagate()
    l1, l2, w = A.PhotonEventGenerator(prim, A.Material("Ar")).simulate_decay_4vectors(100, n_samples=3, seed=7)
    assert isinstance(l1, list) and isinstance(l2, list) and isinstance(w, list) and len(l1) == len(l2) == len(w) == 6
    assert isinstance(l1[0], A.LorentzVector) and (l1[0] + l2[0]).energy() == pytest.approx(100.0, rel=1e-9)
    a1, a2, aw = A.PhotonEventGenerator(prim, A.Material("Ar")).simulate_decay_4vectors(100, n_samples=3, seed=7, strict=False)
    assert np.array_equal(a1.pmu, np.array([v.pmu for v in l1])) and np.array_equal(np.asarray(aw), np.array(w))


def test_compton_rates_statistical(A):
    """Free-running RNG, different seeds: total rate agrees with an independent oracle run within MC error."""
    pf = _c2_flux(200)
    ma, gg, ns = 1.5, 1e-7, 2000
    rates = []
    for seed in (1, 2, 3, 4):
        fl = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=ma, axion_coupling=gg, n_samples=ns, seed=seed, **GEOM)
        fl.simulate(); fl.propagate()
        rates.append(A.ElectronEventGenerator(fl, A.Material("Ar")).decays(100, 5.0))
    us = O.UniformSource(np.random.RandomState(12345))
    oE, oF = O.compton_simulate(pf, ma, gg, 74, O.AbsCrossSection(O.Material("W")), ns, us)
    p = O.propagate(oE, oF, ma, O.W_ee(gg, ma), 1.0, 4.0, 0.2, geom_accept=GA)
    ew, orate = O.decays(oE, p["decay_w"], 100, 5.0)
    sigma = np.sqrt(np.sum(ew.astype(np.float64) ** 2))             # MC standard error of a sum of weights
    assert abs(np.mean(rates) - orate) < 5 * sigma * np.sqrt(1 + 1 / 4)
    assert np.std(rates) < 5 * sigma
    # binned spectrum
    edges = np.linspace(0, 1000, 21)
    fl.propagate()
    import torch
    from alplib_b200.reduce import histogram
    dev = fl.device_arrays()                                         # (launches whatever is still pending)
    h = histogram(dev["axion_energy"], dev["decay_axion_weight"], edges).cpu().numpy()
    oh = O.histogram(oE, p["decay_w"], edges)
    err = np.sqrt(O.histogram(oE, p["decay_w"].astype(np.float64) ** 2, edges))
    sel = oh > 0
    assert sel.sum() >= 10
    assert np.all(np.abs(h[sel] - oh[sel]) < 6 * np.sqrt(2) * err[sel])


def test_propagate_misc(A):
    g = load_golden("propagate_misc")
    fl = A.FluxComptonIsotropic(np.array([[5.0, 1e9]]), A.Material("W"), axion_mass=0.9, axion_coupling=1e-6,
                                n_samples=8, keep_f64=True, **GEOM)
    fl.axion_energy = list(g["axion_energy"])
    fl.axion_flux = list(g["axion_flux"])
    fl.propagate()                                                  # W_ee = 0 below 2 m_e -> infinite lifetime
    assert np.all(fl.decay_axion_weight == 0)
    assert max_rel(fl.scatter_axion_weight, g["zero_scatter_w32"]) <= F32
    fl.timing_window = float(g["tof_window"])
    A.AxionFlux.propagate(fl, float(g["tof_width"]), float(g["tof_rescale"]), time_of_flight_cut=True)
    assert max_rel(fl.scatter_axion_weight_f64, g["tof_scatter_w64"]) < TOL
    assert_decay_close(fl.decay_axion_weight_f64, g["tof_decay_w64"], g["tof_scatter_w64"], TOL)
    # empty flux: every call is a no-op that returns empty arrays / zero rate
    e = A.FluxComptonIsotropic(np.array([[0.3, 1e9]]), A.Material("W"), axion_mass=1.5, axion_coupling=1e-7, n_samples=4, **GEOM)
    e.simulate(); e.propagate()
    assert len(e.axion_energy) == 0 and len(e.decay_axion_weight) == 0
    assert A.ElectronEventGenerator(e, A.Material("Ar")).decays(100, 5.0) == 0.0


@pytest.mark.parametrize("name", ["katd_decay4", "decay4_boosted", "decay4_heavy"])
def test_decay_4vectors_injected(A, name):
    g = load_golden(name)
    ma, ns = float(g["ma"]), int(g["n_samples"])
    fl = A.FluxPrimakoffIsotropic(g["photon_flux"], A.Material("W"), axion_mass=ma, axion_coupling=float(g["g"]), **GEOM)
    fl.simulate(); fl.propagate()
    assert max_rel(fl.decay_axion_weight, g["decay_w32"]) <= F32
    fl.decay_axion_weight = g["decay_w32"]                          # pin the parents' float32 weights to the fixture
    gen = A.PhotonEventGenerator(fl, A.Material("Ar"))
    npar = len(fl.axion_energy)
    u = g["uniforms"][npar:].reshape(npar, 2, ns)                   # first n_flux draws: unused parent phis
    l1, l2, w = gen.simulate_decay_4vectors(int(g["days"]), n_samples=ns, uniforms=u)
    assert len(l1) == len(l2) == len(w) == npar * ns
    scale = np.repeat(np.asarray(fl.axion_energy), ns)[:, None]
    assert np.max(np.abs(l1.pmu - g["p1"]) / scale) < TOL           # |dp| / E_parent
    assert np.max(np.abs(l2.pmu - g["p2"]) / scale) < TOL
    assert np.array_equal(np.asarray(w), g["w"])
    lv = l1[0]
    assert isinstance(lv, A.LorentzVector) and lv.energy() == l1.pmu[0, 0]
    if name == "katd_decay4":
        assert lv.energy() == pytest.approx(46.15041488522627, rel=1e-12)
        assert w[0] == 94.71786499023438 and w[3] == 26.58319664001465


def test_decay_4vectors_free_running_and_invariants(A):
    pf = np.stack([np.logspace(0, 5, 64), np.full(64, 1e12)], axis=1)
    ma, ns, seed = 0.1, 3001, 991
    fl = A.FluxPrimakoffIsotropic(pf, A.Material("W"), axion_mass=ma, axion_coupling=1e-5, seed=seed, **GEOM)
    fl.simulate(); fl.propagate()
    gen = A.PhotonEventGenerator(fl, A.Material("Ar"))
    l1, l2, w = gen.simulate_decay_4vectors(100, n_samples=ns)
    E = np.asarray(fl.axion_energy); npar = E.shape[0]
    # sample by sample against the oracle on the replayed Philox variates
    us = [np.zeros(npar)]
    for r in range(npar):
        u1, u2 = PX.uniform_2(seed, PX.TAG_DECAY2, r, ns)
        us += [u1, u2]
    o1, o2, ow = O.decay_4vectors(E, fl.decay_axion_weight, ma, 100, ns, O.UniformSource(np.concatenate(us)))
    scale = np.repeat(E, ns)[:, None]
    assert np.max(np.abs(l1.pmu - o1) / scale) < TOL and np.max(np.abs(l2.pmu - o2) / scale) < TOL
    assert np.array_equal(np.asarray(w), ow)
    # size-independent invariants: 4-momentum conservation, back-to-back transverse momenta, massless photons.
    # (The reference's parent mass sqrt(E^2-p^2) and gamma = 1/sqrt(1-beta^2) cancel catastrophically for gamma >~ 1e4,
    #  SURVEY section 7 hard part 1, so the invariants are asserted on the moderately boosted parents.)
    p1, p2 = l1.pmu, l2.pmu
    sel = np.repeat(E <= 50.0, ns)
    assert sel.any() and not sel.all()
    pa = np.repeat(np.sqrt(E ** 2 - ma ** 2), ns)
    Ep = scale[:, 0]
    assert np.max(np.abs(p1[sel, 0] + p2[sel, 0] - Ep[sel]) / Ep[sel]) < TOL
    assert np.max(np.abs(p1[sel, 3] + p2[sel, 3] - pa[sel]) / Ep[sel]) < TOL
    assert np.array_equal(p1[:, 1], -p2[:, 1]) and np.array_equal(p1[:, 2], -p2[:, 2])
    m2 = p1[:, 0] ** 2 - p1[:, 1] ** 2 - p1[:, 2] ** 2 - p1[:, 3] ** 2
    assert np.max(np.abs(m2[sel]) / Ep[sel] ** 2) < 1e-9
    assert np.sum(np.asarray(w)) == pytest.approx(100 * 86400 * np.sum(fl.decay_axion_weight.astype(np.float64)), rel=1e-6)


def test_histogram(A):
    import torch
    from alplib_b200.reduce import histogram, total
    rs = np.random.RandomState(3)
    x = np.concatenate([rs.uniform(-1, 11, 200001), [0.0, 10.0, 5.0, np.nan]])
    w = rs.uniform(0, 2, x.shape[0])
    for edges in (np.linspace(0, 10, 41), np.array([0.0, 0.1, 1.0, 2.5, 10.0]), np.logspace(-2, 1, 5000)):
        dx = torch.from_numpy(x).cuda(); dw = torch.from_numpy(w).cuda()
        h = histogram(dx, dw, edges).cpu().numpy()
        ref = O.histogram(x, w, edges)
        assert max_rel(h, ref, floor=1e-9) < 1e-10
        h32 = histogram(dx, dw.float(), edges).cpu().numpy()
        assert max_rel(h32, O.histogram(x, w.astype(np.float32).astype(np.float64), edges), floor=1e-9) < 1e-10
    assert float(total(dw).item()) == pytest.approx(np.sum(w), rel=1e-13)


# ---- e+- channels (BASELINE config 4) -----------------------------------------------------------------------------------
GEOM_DUNE = dict(det_dist=574.0, det_length=5.0, det_area=21.0)
GA_DUNE = O.geom_acceptance(21.0, 574.0)


@pytest.mark.parametrize("name", ["katb_brem", "brem_ps", "brem_vec", "brem_notl"])
def test_brem_injected(A, name):
    g = load_golden(name)
    ma, gg, ns = float(g["ma"]), float(g["g"]), int(g["n_samples"])
    boson = str(g["boson_type"]) if "boson_type" in g else "pseudoscalar"
    tl = bool(g["use_track_length"]) if "use_track_length" in g else True
    fl = A.FluxBremIsotropic(g["electron_flux"], g["positron_flux"], A.Material("C"), target_length=100, axion_mass=ma,
                             axion_coupling=gg, n_samples=ns, boson_type=boson, keep_f64=True, **GEOM_DUNE)
    fl.simulate(use_track_length=tl, uniforms=g["uniforms"])
    assert fl._emit.cpu().numpy().astype(bool).tolist() == g["row_emit"].astype(bool).tolist()
    assert max_rel(fl.axion_energy, g["axion_energy"]) < 1e-12
    assert max_rel(fl.axion_flux, g["axion_flux"]) < TOL
    fl.propagate()
    assert max_rel(fl.scatter_axion_weight_f64 * GA_DUNE, g["scatter_w64"]) < TOL
    assert_decay_close(fl.decay_axion_weight_f64 * GA_DUNE, g["decay_w64"], g["scatter_w64"], TOL)
    assert max_rel(fl.scatter_axion_weight, g["scatter_w32"]) <= F32
    assert_decay_close(fl.decay_axion_weight, g["decay_w32"], g["scatter_w32"], TOL, extra_rel=F32)


def _brem_tables(n):
    e = np.linspace(20.0, 1.2e5, n)
    return np.stack([e, 1e14 * np.exp(-e / 2e4)], axis=1), np.stack([e, 0.3e14 * np.exp(-e / 2e4)], axis=1)


def test_brem_free_running_matches_oracle(A):
    ef, pfx = _brem_tables(53)
    ef[0, 0] = 5.0                                        # below ep_min = ma: filtered by the host
    ma, gg, ns, seed = 10.0, 1e-6, 301, 4242
    fl = A.FluxBremIsotropic(ef, pfx, A.Material("C"), target_length=100, axion_mass=ma, axion_coupling=gg, n_samples=ns,
                             seed=seed, **GEOM_DUNE)
    fl.simulate()
    # replay the kernel's Philox variates in the reference's draw order
    stream = []
    for r in range(ef.shape[0]):
        if ef[r, 0] < max(ma, O.M_E):
            continue
        u = PX.uniform_1(seed, PX.TAG_BREM_ROW, r, 1)[0]
        e1 = max(ma, O.M_E) + (ef[r, 0] - max(ma, O.M_E)) * u
        stream.append([u])
        if e1 * (1 - np.power(ma / e1, 2)) > ma:
            stream.append(PX.uniform_1(seed, PX.TAG_BREM_EA, r, ns))
    oE, oF, info = O.brem_simulate(ef, pfx, ma, gg, O.Material("C"), ns, O.UniformSource(np.concatenate(stream)))
    assert len(fl.axion_energy) == oE.shape[0] > 0
    assert max_rel(fl.axion_energy, oE) < 1e-12
    assert max_rel(fl.axion_flux, oF) < TOL
    fl.propagate()
    p = O.propagate(oE, oF, ma, O.W_ee(gg, ma), 1.0, 574.0, 5.0, geom_accept=GA_DUNE)
    assert max_rel(fl.scatter_axion_weight, p["scatter_w"]) <= F32
    r_gpu = A.ElectronEventGenerator(fl, A.Material("Ar")).decays(365, 30.0)
    _, r_cpu = O.decays(oE, p["decay_w"], 365, 30.0)
    assert r_gpu == pytest.approx(r_cpu, rel=1e-5)
    # NA64-style single-row monoenergetic beam (examples/ex4_na64.py shape)
    beam = np.array([[1e5, 2.84e11 / 86400 / 365]])
    mono = A.FluxBremIsotropic(beam, target=A.Material("W"), det_dist=43.3, det_length=1.0, det_area=0.23,
                               target_length=2.0, axion_mass=10.0, axion_coupling=1e-5, n_samples=1000,
                               is_isotropic=False, is_monoenergetic=True, seed=3)
    mono.simulate(); mono.propagate()
    assert len(mono.axion_energy) == 1000 and np.all(np.isfinite(mono.axion_flux)) and np.sum(mono.axion_flux) > 0
    with pytest.raises(TypeError):
        A.FluxBremIsotropic(beam, target=A.Material("W"), axion_mass=10.0).simulate()     # default 1-D positron_flux


def test_resonance(A):
    g = load_golden("katr_resonance")
    fl = A.FluxResonanceIsotropic(g["positron_flux"], A.Material("C"), axion_mass=10.0, axion_coupling=1e-6,
                                  n_samples=1000, keep_f64=True, **GEOM_DUNE)
    fl.simulate(uniforms=g["uniforms"].reshape(2, 1000))
    assert list(fl.axion_energy) == [97.84735812133073]
    assert max_rel(fl.axion_flux, g["axion_flux"]) < TOL
    fl.propagate()
    assert max_rel(fl.scatter_axion_weight, g["scatter_w32"]) <= F32
    assert_decay_close(fl.decay_axion_weight, g["decay_w32"], g["scatter_w32"], TOL, extra_rel=F32)
    # free-running: replay Philox through the oracle
    seed, ns = 99, 20001
    fr = A.FluxResonanceIsotropic(g["positron_flux"], A.Material("C"), axion_mass=10.0, axion_coupling=1e-6,
                                  n_samples=ns, seed=seed, **GEOM_DUNE)
    fr.simulate()
    u1, u2 = PX.uniform_2(seed, PX.TAG_RESONANCE, 0, ns)
    oE, oF = O.resonance_simulate(g["positron_flux"], 10.0, 1e-6, O.Material("C"), ns, O.UniformSource(np.concatenate([u1, u2])))
    assert max_rel(fr.axion_flux, oF) < TOL
    # guards: below 2 m_e / above the flux table -> no sample (fluxes.py:428-435)
    for ma in (0.9, 200.0):
        e = A.FluxResonanceIsotropic(g["positron_flux"], A.Material("C"), axion_mass=ma, axion_coupling=1e-6, **GEOM_DUNE)
        e.simulate(); e.propagate()
        assert len(e.axion_energy) == 0 and len(e.decay_axion_weight) == 0


@pytest.mark.parametrize("name", ["kata_pairann", "pairann_mid"])
def test_pairann_injected(A, name):
    g = load_golden(name)
    ns = int(g["n_samples"])
    fl = A.FluxPairAnnihilationIsotropic(g["positron_flux"], A.Material("C"), axion_mass=10.0, axion_coupling=1e-6,
                                         n_samples=ns, keep_f64=True, **GEOM_DUNE)
    fl.simulate(uniforms=g["uniforms"])
    assert max_rel(fl.axion_energy, g["axion_energy"]) < TOL
    assert max_rel(fl.axion_flux, g["axion_flux"]) < TOL
    fl.propagate()
    assert max_rel(fl.scatter_axion_weight_f64 * GA_DUNE, g["scatter_w64"]) < TOL
    assert_decay_close(fl.decay_axion_weight_f64 * GA_DUNE, g["decay_w64"], g["scatter_w64"], TOL)
    assert max_rel(fl.scatter_axion_weight, g["scatter_w32"]) <= F32


def test_pairann_free_running_matches_oracle(A):
    _, pfx = _brem_tables(41)
    ma, gg, ns, seed = 10.0, 1e-6, 257, 777
    fl = A.FluxPairAnnihilationIsotropic(pfx, A.Material("C"), axion_mass=ma, axion_coupling=gg, n_samples=ns, seed=seed,
                                         **GEOM_DUNE)
    fl.simulate()
    keep = O.pairann_row_valid(pfx, ma)
    assert 0 < keep.sum() < pfx.shape[0]
    stream = []
    for r in np.nonzero(keep)[0]:
        stream += [np.zeros(ns), PX.uniform_1(seed, PX.TAG_PAIRANN, int(r), ns)]     # phi draws do not enter E_lab
    oE, oF = O.pairann_simulate(pfx, ma, gg, O.Material("C"), ns, O.UniformSource(np.concatenate(stream)))
    assert max_rel(fl.axion_energy, oE) < TOL
    assert max_rel(fl.axion_flux, oF) < TOL
    # fused one-kernel chain == separate launches
    fl.propagate()
    r_sep = A.ElectronEventGenerator(fl, A.Material("Ar")).decays(365, 30.0)
    fb = A.FluxPairAnnihilationIsotropic(pfx, A.Material("C"), axion_mass=ma, axion_coupling=gg, n_samples=ns, seed=seed,
                                         fused=True, **GEOM_DUNE)
    fb.simulate(); fb.propagate()
    assert A.ElectronEventGenerator(fb, A.Material("Ar")).decays(365, 30.0) == pytest.approx(r_sep, rel=1e-12)


# ---- batched (m_a, g) scan (BASELINE config 5) ------------------------------------------------------------------------------
@pytest.mark.parametrize("process", ["compton", "primakoff"])
def test_scan_matches_per_point_api(A, process):
    from alplib_b200.scan import scan_rates
    pf = _c2_flux(60)
    ma_grid = np.array([0.01, 0.3, 1.5, 4.0, 30.0, 2000.0])          # last mass: above every photon energy -> all zero
    g_grid = np.logspace(-8, -4, 7) if process == "compton" else np.logspace(-7, -3, 7)
    ns, seed = 128, 31337
    rates = scan_rates(process, pf, A.Material("W"), ma_grid, g_grid, n_samples=ns, days_exposure=100, threshold=5.0,
                       seed=seed, **GEOM)
    assert rates.shape == (6, 7) and np.all(rates[-1] == 0) and np.all(np.isfinite(rates))
    ref = np.zeros_like(rates)
    for i, m in enumerate(ma_grid):
        if process == "compton":
            fl = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=m, axion_coupling=g_grid[0], n_samples=ns, seed=seed, **GEOM)
            gen = A.ElectronEventGenerator(fl, A.Material("Ar"))
        else:
            fl = A.FluxPrimakoffIsotropic(pf, A.Material("W"), axion_mass=m, axion_coupling=g_grid[0], **GEOM)
            gen = A.PhotonEventGenerator(fl, A.Material("Ar"))
        fl.simulate()
        for j, g in enumerate(g_grid):
            fl.propagate(new_coupling=g)
            ref[i, j] = gen.decays(100, 5.0)
    nz = ref != 0
    assert nz.sum() > 15
    assert max_rel(rates[nz], ref[nz]) < 1e-11                        # same samples, same arithmetic: only the sum order differs
    assert np.all(rates[~nz] == 0)


def test_scan_matches_oracle(A):
    from alplib_b200.scan import scan_rates
    pf = _c2_flux(25)
    ma_grid, g_grid, ns, seed = np.array([0.5, 2.0]), np.array([1e-7, 3e-6, 1e-4]), 200, 5
    rates = scan_rates("compton", pf, A.Material("W"), ma_grid, g_grid, n_samples=ns, days_exposure=100, threshold=5.0,
                       seed=seed, **GEOM)
    W = O.AbsCrossSection(O.Material("W"))
    for i, m in enumerate(ma_grid):
        keep = O.compton_row_valid(pf, m)
        u = np.concatenate([PX.uniform_1(seed, PX.TAG_COMPTON_EA, int(r), ns) for r in np.nonzero(keep)[0]])
        oE, oF = O.compton_simulate(pf, m, g_grid[0], 74, W, ns, O.UniformSource(u))
        for j, g in enumerate(g_grid):
            p = O.propagate(oE, oF, m, O.W_ee(g, m), np.power(g / g_grid[0], 2), 4.0, 0.2, geom_accept=GA)
            ew, r = O.decays(oE, p["decay_w"], 100, 5.0)
            # decay_prob = 1-exp(-x) cancels for long-lived points: compare with the summed conditioned tolerance
            tol = float(np.sum(np.abs(ew) * (decay_tol(p["decay_w64"], p["scatter_w64"], 1e-9) + 2 * F32)) / max(abs(r), 1e-300))
            assert abs(rates[i, j] - r) <= tol * abs(r) + 1e-300, (i, j, rates[i, j], r, tol)


# ---- propagate_iso_vol_int + DetectorGeometry -------------------------------------------------------
@pytest.mark.parametrize("name", ["volint_box", "volint_sphere", "volint_cyl"])
def test_propagate_iso_vol_int(A, name):
    g = load_golden(name)
    ma, gg, nc = float(g["ma"]), float(g["g"]), float(g["new_coupling"])
    fl = A.FluxComptonIsotropic(g["photon_flux"], A.Material("W"), axion_mass=ma, axion_coupling=gg, n_samples=12,
                                keep_f64=True, **GEOM)
    fl.axion_energy = g["axion_energy"]
    fl.axion_flux = g["axion_flux"]
    np.random.seed(78)                                   # the fixture's geometry draws (legacy global RNG, as the reference)
    geom = A.DetectorGeometry(float(g["l0"]), str(g["geometry"]), list(g["params"]), n_samples=int(g["n_geom"]))
    fl.propagate_iso_vol_int(geom, new_coupling=nc)
    assert max_rel(fl.decay_axion_weight_f64, g["decay_w64"]) < TOL
    assert max_rel(fl.scatter_axion_weight_f64, g["scatter_w64"]) < TOL
    assert max_rel(fl.decay_axion_weight, g["decay_w32"]) <= F32 and max_rel(fl.scatter_axion_weight, g["scatter_w32"]) <= F32
    assert fl.decay_axion_weight.dtype == np.float32
    rate = A.ElectronEventGenerator(fl, A.Material("Ar")).decays(100, 1.0)
    _, orate = O.decays(g["axion_energy"], g["decay_w32"], 100, 1.0)
    assert rate == pytest.approx(orate, rel=1e-6)
    with pytest.raises(Exception, match="closer than 0.0"):
        A.DetectorGeometry(0.0)


def test_propagate_iso_vol_int_large(A):
    """1e4 samples x 1e4 volume points (the reference needs ~1 s per 100 samples): long-detector limit check --
    for a thin, distant box the volume integral tends to decay_prob * area/(4 pi L^2) of the plain propagate."""
    pf = _c2_flux(100)
    fl = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=1.5, axion_coupling=1e-6, n_samples=100, seed=4, keep_f64=True,
                                det_dist=50.0, det_length=0.2, det_area=0.04)
    fl.simulate()
    np.random.seed(1)
    geom = A.DetectorGeometry(50.0, "box", [0.2, 0.2, 0.2], n_samples=10000)
    fl.propagate_iso_vol_int(geom)
    vol = np.array(fl.decay_axion_weight_f64)
    fl.propagate()
    plain = np.array(fl.decay_axion_weight_f64) * O.geom_acceptance(0.04, 50.0)
    sel = plain > 1e-3 * plain.max()
    # same physics up to O(det_length/det_dist) geometry corrections and the MC error of 1e4 points
    assert np.max(np.abs(vol[sel] / plain[sel] - 1)) < 0.03


@pytest.mark.parametrize("hist_rows", ["1", "0"])         # fused histogram in the row kernel / in the tile kernel
@pytest.mark.parametrize("ns", [999, 1000, 20000])      # odd: generic kernel; even: paired kernel, short and long rows
def test_chain_fused_histogram(A, ns, hist_rows, monkeypatch):
    """Event-weighted E_a spectrum accumulated inside the chain kernel == np.histogram of the materialised arrays: linear,
    logarithmic and irregular edges (spacing guess on / off), few bins (row kernel: private bin columns per warp, 32 / 16
    columns; tile kernel: one copy per warp, aggregated adds) and many (one copy per CTA, shared atomics)."""
    from alplib_b200.reduce import histogram
    monkeypatch.setenv("ALPK_HIST_ROWS", hist_rows)
    pf = _c2_flux(150 if ns < 20000 else 40)
    fl = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=1.5, axion_coupling=3e-6, n_samples=ns, seed=8, **GEOM)
    fl.simulate()
    irregular = np.array([0.0, 1.5, 1.6, 2.0, 7.0, 7.5, 40.0, 300.0, 301.0, 1000.0])
    for edges in (np.linspace(0.0, 1000.0, 51), np.logspace(0, 3, 33), irregular, np.logspace(0, 3, 201), np.logspace(0, 3, 2001)):
        rate, evw = fl.chain(100, 5.0, materialize=True, hist_edges=edges)
        h = fl.last_hist.cpu().numpy()
        ref = O.histogram(np.asarray(fl.axion_energy), evw.cpu().numpy(), edges)
        assert max_rel(h, ref, floor=1e-300) < 1e-10
        assert max_rel(histogram(fl._E, evw, edges).cpu().numpy(), ref, floor=1e-300) < 1e-10
        assert h.sum() == pytest.approx(float(rate.item()) - evw.cpu().numpy()[np.asarray(fl.axion_energy) > edges[-1]].sum()
                                        - evw.cpu().numpy()[np.asarray(fl.axion_energy) < edges[0]].sum(), rel=1e-9)
        # reduce-only pass fills the same histogram without materialising anything
        rate2, _ = fl.chain(100, 5.0, materialize=False, hist_edges=edges)
        assert max_rel(fl.last_hist.cpu().numpy(), ref, floor=1e-300) < 1e-10


def test_output_bounds_canaries(A):
    """compute-sanitizer is closed on this pool: check for out-of-bounds stores with guard regions instead.  Every output
    of the chain / decay / propagate kernels is a slice of a sentinel-filled buffer; the guards must be untouched."""
    import ctypes as C
    import torch
    from alplib_b200 import _native as N, params as P
    pf = _c2_flux(23)
    fl = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=1.5, axion_coupling=1e-6, n_samples=333, seed=1, **GEOM)   # odd n
    fl._stage_inputs(); fl._launch_rows()
    n = fl._rows[1] * fl._rows[2]
    assert n % 2 == 1
    G, S = 64, -7.25
    rt = fl._rt

    def guarded(count, dtype):
        buf = torch.full((count + 2 * G,), S, dtype=dtype, device=rt.device)
        return buf, buf[G:G + count]

    bufs = {k: guarded(n, torch.float64) for k in ("E", "F", "d64", "s64", "evw")}
    bufs.update({k: guarded(n, torch.float32) for k in ("d32", "s32")})
    rate = rt.empty(1)
    prop = fl._prop_params(*fl._width_rescale(None), fl._geom_accept(), None)
    fl._launch_samples(bufs["E"][1], bufs["F"][1], prop, P.event_params(100, 5.0), d32=bufs["d32"][1], s32=bufs["s32"][1],
                       d64=bufs["d64"][1], s64=bufs["s64"][1], evw=bufs["evw"][1], rate=rate)
    torch.cuda.synchronize()
    for k, (buf, view) in bufs.items():
        assert torch.all(buf[:G] == S) and torch.all(buf[-G:] == S), k
        assert not torch.any(view == S), k
    # decay MC with an odd event count (odd planes are not 16-byte aligned)
    fp = A.FluxPrimakoffIsotropic(np.array([[100.0, 1e12], [40.0, 1e11], [7.0, 1e10]]), A.Material("W"), axion_mass=0.1,
                                  axion_coupling=1e-5, **GEOM)
    fp.simulate(); fp.propagate(); fp.device_arrays()
    npar, ns = 3, 111
    nev = npar * ns
    table = rt.empty(npar * N.PARENT_NP)
    ev = P.event_params(100, 0.0)
    N.call("alpk_decay_parent_rows", N.ptr(fp._E), N.ptr(fp._D32), npar, float(0.1 ** 2), C.byref(ev), ns, N.ptr(table), rt.stream())
    b1, p1 = guarded(4 * nev, torch.float64); b2, p2 = guarded(4 * nev, torch.float64); bw, w = guarded(nev, torch.float64)
    for fast in (0, 1):
        N.call("alpk_decay2body", N.ptr(table), npar, ns, None, None, C.c_uint64(5), fast, N.ptr(p1), N.ptr(p2), N.ptr(w), rt.stream())
        torch.cuda.synchronize()
        for buf, view in ((b1, p1), (b2, p2), (bw, w)):
            assert torch.all(buf[:G] == S) and torch.all(buf[-G:] == S) and not torch.any(view == S)


# ---- BASELINE full sizes: size-independent properties, evaluated on the device ------------------------------------------------
def test_full_size_c2_chain_properties(A):
    """configs[1] at full size (1e4 bins x 1e4 samples = 7.4e7 live samples): the fused chain's invariants."""
    import torch
    pf = _c2_flux(10_000)
    ma = 1.5
    fl = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=ma, axion_coupling=1e-7, n_samples=10_000, seed=0xA1B2C3D4 + 2, **GEOM)
    fl.simulate_rows_only = None
    fl._reset_samples(); fl._stage_inputs(); fl._launch_rows()
    edges = np.concatenate([[0.0], np.logspace(0, 3, 60), [1e9]])
    rate, evw = fl.chain(100, 5.0, materialize=True, hist_edges=edges)
    n_live = int(O.compton_row_valid(pf, ma).sum()) * 10_000
    E, F, D, S = fl._E, fl._F, fl._D32, fl._S32
    assert E.shape[0] == n_live == evw.shape[0] and 7.3e7 < n_live < 7.5e7
    r = float(rate.item())
    assert r == pytest.approx(float(evw.sum().item()), rel=1e-10)
    assert float(fl.last_hist.sum().item()) == pytest.approx(r, rel=1e-10)           # checksum of the binned spectrum
    rate2, _ = fl.chain(100, 5.0, materialize=False)
    assert float(rate2.item()) == pytest.approx(r, rel=1e-12)                         # reduce-only == materialising
    assert bool(torch.isfinite(F).all()) and bool(torch.isfinite(D).all()) and bool((F >= 0).all())
    assert float(E.min().item()) >= ma and float(E.max().item()) <= pf[-1, 0]
    assert bool((D <= S).all()) and bool((D >= 0).all())                              # decay prob <= 1
    eg = torch.from_numpy(np.repeat(pf[O.compton_row_valid(pf, ma), 0], 10_000)).to(E.device)
    assert bool((E <= eg).all())                                                      # E_a drawn in [ma, E_gamma] of its own row
    # the eager three-kernel path reproduces the same arrays bit for bit
    fl2 = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=ma, axion_coupling=1e-7, n_samples=10_000, seed=0xA1B2C3D4 + 2,
                                 fused=False, **GEOM)
    fl2.simulate(); fl2.propagate()
    assert torch.equal(fl2._E, E) and torch.equal(fl2._F, F) and torch.equal(fl2._D32, D) and torch.equal(fl2._S32, S)


def test_full_size_c3_decay_properties(A):
    """configs[2]: 1e4 parents x 1e4 decays = 1e8 events (7.2 GB of 4-vectors): conservation laws on the device."""
    import torch
    pf = _c2_flux(10_000)
    fl = A.FluxPrimakoffIsotropic(pf, A.Material("W"), axion_mass=0.1, axion_coupling=1e-5, **GEOM)
    fl.simulate(); fl.propagate()
    gen = A.PhotonEventGenerator(fl, A.Material("Ar"))
    l1, l2, w = gen.simulate_decay_4vectors(100, n_samples=10_000, seed=3)
    p1, p2 = l1.device, l2.device
    n = p1.shape[1]
    assert n == 10_000 * 10_000
    Ep = fl._E.repeat_interleave(10_000)
    assert torch.equal(p1[1], -p2[1]) and torch.equal(p1[2], -p2[2])                  # back-to-back transverse momenta
    # the reference's sqrt(E^2-p^2) / 1/sqrt(1-beta^2) lose ~gamma^2 * 1e-16; assert where that is < 1e-9 (gamma <= 1000)
    sel = Ep <= 100.0
    dE = ((p1[0] + p2[0] - Ep).abs() / Ep)[sel]
    assert float(dE.max().item()) < 1e-9
    pz = torch.sqrt(Ep * Ep - 0.1 ** 2)
    assert float((((p1[3] + p2[3] - pz).abs() / Ep)[sel]).max().item()) < 1e-9
    m2 = (p1[0] ** 2 - p1[1] ** 2 - p1[2] ** 2 - p1[3] ** 2).abs() / Ep ** 2
    assert float(m2[sel].max().item()) < 1e-9
    assert bool((p1[0] > 0).all()) and bool((p2[0] > 0).all())
    assert float(w.device.sum().item()) == pytest.approx(100 * 86400 * float(fl._D32.double().sum().item()), rel=1e-6)
    # isotropy in the parent frame <=> E1/E_parent uniform in [(1-beta)/2, (1+beta)/2]: mean 1/2, variance beta^2/12
    x = (p1[0] / Ep)[~sel]
    assert abs(float(x.mean().item()) - 0.5) < 1e-3 and abs(float(x.var().item()) - 1 / 12) < 1e-3


def test_c_abi_without_torch_tensors(A):
    """The C ABI alone (ctypes + the library's own device-memory helpers, no torch tensors): README Primakoff chain."""
    import ctypes as C
    from alplib_b200 import _native as N, params as P
    L = N.lib()
    g = load_golden("c1_primakoff")
    xs = O.AbsCrossSection(O.Material("W"))
    host = {"eg": np.array([100.0]), "ph": np.array([1e12]), "le": np.ascontiguousarray(xs.log10_e), "lx": np.ascontiguousarray(xs.log10_xs)}
    dev = {}

    def dmalloc(nbytes):
        p = C.c_void_p()
        assert L.alpk_malloc(C.byref(p), nbytes) == 0, L.alpk_last_error()
        return p

    assert L.alpk_set_device(0) == 0
    for k, a in host.items():
        dev[k] = dmalloc(a.nbytes)
        assert L.alpk_memcpy_h2d(dev[k], a.ctypes.data_as(C.c_void_p), a.nbytes, None) == 0
    for k, nb in (("E", 8), ("F", 8), ("d32", 4), ("s32", 4), ("rate", 8), ("scratch", 8 * N.SCRATCH_DOUBLES)):
        dev[k] = dmalloc(nb)
    pref, r02 = P.primakoff_consts(1e-5, np.int64(74))
    assert L.alpk_primakoff(dev["eg"], dev["ph"], 1, dev["le"], dev["lx"], host["le"].shape[0], 0.1, pref, r02, dev["E"], dev["F"], None) == 0
    prop = P.prop_params(0.1, O.W_gg(1e-5, 0.1), 1.0, 4.0, 0.2, geom_accept=GA, fast=False)
    assert L.alpk_propagate(dev["E"], dev["F"], 1, C.byref(prop), dev["d32"], dev["s32"], None, None, None) == 0
    ev = P.event_params(100, 5.0)
    assert L.alpk_decays(dev["E"], dev["d32"], 1, C.byref(ev), None, dev["rate"], dev["scratch"], None) == 0
    rate, flux, d32 = np.zeros(1), np.zeros(1), np.zeros(1, np.float32)
    for dst, src in ((rate, "rate"), (flux, "F"), (d32, "d32")):
        assert L.alpk_memcpy_d2h(dst.ctypes.data_as(C.c_void_p), dev[src], dst.nbytes, None) == 0
    assert L.alpk_sync(None) == 0
    assert flux[0] == pytest.approx(327952.3965351194, rel=1e-12)
    assert d32[0] == g["decay_w32"][0] and rate[0] == pytest.approx(284.1535949707031, rel=2e-7)
    # error path: status code + message, no exception across the boundary
    assert L.alpk_propagate(None, None, 5, None, None, None, None, None, None) != 0
    assert b"alpk_propagate" in L.alpk_last_error()
    for p in dev.values():
        assert L.alpk_free(p) == 0


# ---- further flux sources ----------------------------------------------------------------------------
def test_nuclear_flux(A):
    g = load_golden("nuclear")
    fl = A.FluxNuclearIsotropic(g["rates"], A.Material("W"), axion_mass=0.3, gann0=float(g["gann0"]), gann1=float(g["gann1"]),
                                gae=float(g["gae"]), keep_f64=True, **GEOM)
    fl.simulate(j=1, delta=0.2)
    assert np.array_equal(fl.axion_energy, g["axion_energy"])
    assert max_rel(fl.axion_flux, g["axion_flux"]) < TOL
    fl.propagate()
    assert max_rel(fl.scatter_axion_weight_f64 * GA, g["scatter_w64"]) < TOL
    assert_decay_close(fl.decay_axion_weight_f64 * GA, g["decay_w64"], g["scatter_w64"], TOL)
    assert max_rel(fl.scatter_axion_weight, g["scatter_w32"]) <= F32
    assert A.ElectronEventGenerator(fl, A.Material("Ar")).decays(100, 0.1) >= 0.0


def test_pi0_flux(A):
    g = load_golden("pi0_flux")
    fp = A.FluxPi0Isotropic(pi0_rate=0.0259, boson_mass=30.0, coupling=1e-3, det_dist=20.0, det_area=2.0, det_length=1.0,
                            target=A.Material("W"))
    fp.simulate_flux(g["momenta"], energy_cut=50.0, angle_cut=0.3, uniforms=g["uniforms"])
    assert max_rel(fp.axion_energy, g["axion_energy"]) < TOL
    assert max_rel(fp.axion_flux, g["axion_flux"]) < TOL
    fp.propagate()
    assert fp.scatter_axion_weight.dtype == np.float64                 # this class keeps float64 weights
    assert max_rel(fp.scatter_axion_weight, g["scatter_w"]) < 1e-15 and np.all(fp.decay_axion_weight == 0)
    with pytest.raises(TypeError):
        A.PhotonEventGenerator(fp, A.Material("Ar")).decays(100, 0.0)
    r = load_golden("pi0_rest")
    fr = A.FluxPi0Isotropic(pi0_rate=0.0259, boson_mass=30.0, coupling=1e-3, n_samples=5, target=A.Material("W"))
    fr.simulate(); fr.propagate(is_isotropic=False)
    assert max_rel(fr.axion_energy, r["axion_energy"]) < 1e-15 and max_rel(fr.scatter_axion_weight, r["scatter_w"]) < 1e-15
    heavy = A.FluxPi0Isotropic(boson_mass=500.0, target=A.Material("W"))
    heavy.simulate()
    assert len(heavy.axion_energy) == 0                               # kinematically forbidden (fluxes.py:974-976)


def test_pair_annihilation_gamma(A):
    g = load_golden("pairgamma")
    ns = int(g["n_samples"])
    fl = A.FluxPairAnnihilationGamma(g["positron_flux"], A.Material("C"), target_radiation_length=float(g["rad_length"]),
                                     axion_mass=float(g["ma"]), axion_coupling=float(g["g"]), n_samples=ns, keep_f64=True, **GEOM_DUNE)
    fl.simulate(uniforms=g["uniforms"])
    assert max_rel(fl.axion_energy, g["axion_energy"]) < TOL
    assert max_rel(fl.axion_flux, g["axion_flux"]) < TOL
    fl.propagate()
    assert max_rel(fl.scatter_axion_weight_f64 * GA_DUNE, g["scatter_w64"]) < TOL
    assert_decay_close(fl.decay_axion_weight_f64 * GA_DUNE, g["decay_w64"], g["scatter_w64"], TOL)
    assert max_rel(fl.scatter_axion_weight, g["scatter_w32"]) <= F32
    # free-running mode: finite, positive, reproducible with the seed
    a = A.FluxPairAnnihilationGamma(g["positron_flux"], A.Material("C"), axion_mass=8.0, axion_coupling=1e-5, n_samples=500, seed=5, **GEOM_DUNE)
    b = A.FluxPairAnnihilationGamma(g["positron_flux"], A.Material("C"), axion_mass=8.0, axion_coupling=1e-5, n_samples=500, seed=5, **GEOM_DUNE)
    a.simulate(); b.simulate()
    assert np.array_equal(a.axion_flux, b.axion_flux) and np.all(np.isfinite(a.axion_flux)) and np.sum(a.axion_flux) > 0


@pytest.mark.parametrize("name", ["meson2body", "meson2body_cut"])
def test_neutral_meson_flux(A, name):
    g = load_golden(name)
    fl = A.FluxNeutralMeson2BodyDecay(g["meson_flux"], flux_weight=float(g["flux_weight"]), boson_mass=float(g["ma"]),
                                      coupling=float(g["g"]), det_dist=float(g["det_dist"]), det_area=float(g["det_area"]),
                                      det_length=1.0, apply_angle_cut=bool(g["angle_cut"]), n_samples=int(g["n_samples"]),
                                      off_axis_angle=float(g["off_axis"]))
    assert len(fl.axion_angle) == 0
    fl.simulate(uniforms=g["uniforms"])
    assert max_rel(fl.axion_energy, g["axion_energy"]) < TOL
    assert max_rel(fl.axion_flux, g["axion_flux"]) < 1e-14
    assert max_rel(fl.axion_angle, g["axion_angle"]) < TOL
    fl.propagate()
    assert fl.scatter_axion_weight.dtype == np.float64 and max_rel(fl.scatter_axion_weight, g["scatter_w"]) < 1e-14
    n0 = len(fl.axion_energy)
    fl.simulate(uniforms=g["uniforms"])                       # the reference appends (no reset in simulate)
    assert len(fl.axion_energy) == 2 * n0 == len(fl.axion_angle)


def test_decay_4vectors_with_polar_angles(A):
    g = load_golden("decay4_angled")
    ma, ns = float(g["ma"]), int(g["n_samples"])
    fl = A.FluxPrimakoffIsotropic(g["photon_flux"], A.Material("W"), axion_mass=ma, axion_coupling=float(g["g"]), **GEOM)
    fl.simulate(); fl.propagate()
    fl.decay_axion_weight = g["decay_w32"]
    fl.axion_angle = list(g["axion_angle"])
    gen = A.PhotonEventGenerator(fl, A.Material("Ar"))
    l1, l2, w = gen.simulate_decay_4vectors(int(g["days"]), n_samples=ns, uniforms=g["uniforms"].reshape(3, 2, ns),
                                            parent_phis=g["parent_phis"])
    scale = np.repeat(np.asarray(fl.axion_energy), ns)[:, None]
    assert np.max(np.abs(l1.pmu - g["p1"]) / scale) < TOL and np.max(np.abs(l2.pmu - g["p2"]) / scale) < TOL
    assert np.array_equal(np.asarray(w), g["w"])
    # free running: the azimuths come from the global NumPy stream, like the reference
    np.random.seed(19)
    a1, _, _ = gen.simulate_decay_4vectors(100, n_samples=ns, seed=1)
    np.random.seed(19)
    b1, _, _ = gen.simulate_decay_4vectors(100, n_samples=ns, seed=1)
    assert np.array_equal(a1.pmu, b1.pmu)
    tot = a1.pmu  # 4-momentum conservation with the angled parents
    ea = np.asarray(fl.axion_energy)
    assert np.all(np.isfinite(tot)) and abs(np.mean(a1.energy()) / np.mean(ea) - 0.5) < 0.1


# ---- randomised differential test: free-running kernels vs the oracle on replayed Philox variates --------------------------------
def test_compton_chain_fuzz(A):
    rs = np.random.RandomState(2024)
    W = O.AbsCrossSection(O.Material("W"))
    for trial in range(14):
        n_rows = int(rs.choice([1, 2, 7, 40, 130]))
        ns = int(rs.choice([1, 2, 3, 17, 43, 44, 45, 256, 1001]))
        ma = float(10 ** rs.uniform(-2, 1.3))
        g = float(10 ** rs.uniform(-8, -3.5))
        eg = np.sort(10 ** rs.uniform(-1, 4, n_rows))
        pf = np.stack([eg, 10 ** rs.uniform(5, 15, n_rows)], axis=1)
        dist, length, area = float(10 ** rs.uniform(0, 2.5)), float(10 ** rs.uniform(-1, 1)), float(10 ** rs.uniform(-2, 1.5))
        loop, iso = bool(rs.randint(2)), bool(rs.randint(2))
        seed = int(rs.randint(1, 2 ** 31))
        fl = A.FluxComptonIsotropic(pf, A.Material("W"), det_dist=dist, det_length=length, det_area=area, axion_mass=ma,
                                    axion_coupling=g, n_samples=ns, is_isotropic=iso, loop_decay=loop, seed=seed, keep_f64=True)
        fl.row_offset = 11 * trial
        fl.simulate(); fl.propagate()
        keep = O.compton_row_valid(pf, ma)
        if not keep.any():
            assert len(fl.axion_energy) == 0
            continue
        u = np.concatenate([PX.uniform_1(seed, PX.TAG_COMPTON_EA, int(r) + 11 * trial, ns) for r in np.nonzero(keep)[0]])
        oE, oF = O.compton_simulate(pf, ma, g, 74, W, ns, O.UniformSource(u))
        width = O.W_ee(g, ma) + (O.W_gg_loop(g, ma, O.M_E) if loop else 0.0)
        ga = O.geom_acceptance(area, dist) if iso else None
        p = O.propagate(oE, oF, ma, width, 1.0, dist, length, geom_accept=ga)
        ctx = (trial, n_rows, ns, ma, g)
        assert np.array_equal(fl.axion_energy, oE), ctx
        assert max_rel(fl.axion_flux, oF) < TOL, ctx
        assert max_rel(fl.scatter_axion_weight_f64, p["scatter_w64"]) < TOL, ctx
        assert_decay_close(fl.decay_axion_weight_f64, p["decay_w64"], p["scatter_w64"], TOL)
        assert max_rel(fl.scatter_axion_weight, p["scatter_w"], floor=1e-44) <= F32, ctx      # (float32 denormals excluded)
        # fused reduce-only rate == separate launches, and == the oracle within the conditioned tolerance
        thr = float(rs.uniform(0, eg.max()))
        r_sep = A.ElectronEventGenerator(fl, A.Material("Ar")).decays(30, thr)
        fb = A.FluxComptonIsotropic(pf, A.Material("W"), det_dist=dist, det_length=length, det_area=area, axion_mass=ma,
                                    axion_coupling=g, n_samples=ns, is_isotropic=iso, loop_decay=loop, seed=seed, fused=True)
        fb.row_offset = 11 * trial
        fb.simulate(); fb.propagate()
        r_fused = A.ElectronEventGenerator(fb, A.Material("Ar")).decays(30, thr)
        assert r_fused == pytest.approx(r_sep, rel=1e-12, abs=1e-300), ctx


def test_full_size_c4_epm_channels(A):
    """configs[3]: 1e4-row e+- shower fluxes on graphite, DUNE-like geometry, ma = 10 MeV, g = 1e-6:
    Brem 1e4 x 1000, Resonance 1e6 draws, PairAnn 1e4 x 100.  Device-side invariants + a sub-sampled oracle comparison."""
    import torch
    e = np.linspace(20.0, 1.2e5, 10_000)
    ef = np.stack([e, 1e14 * np.exp(-e / 2e4)], axis=1)
    pfx = np.stack([e, 0.3e14 * np.exp(-e / 2e4)], axis=1)
    ma, g = 10.0, 1e-6
    C_ = A.Material("C")
    # --- Brem
    fb = A.FluxBremIsotropic(ef, pfx, C_, target_length=100, axion_mass=ma, axion_coupling=g, n_samples=1000, seed=11, **GEOM_DUNE)
    fb.simulate(); fb.propagate(); fb.device_arrays()
    n_emit = int(fb._emit.sum().item())
    assert fb._E.shape[0] == n_emit * 1000 and n_emit > 9000
    assert bool(torch.isfinite(fb._F).all()) and bool((fb._F >= 0).all()) and float(fb._E.min().item()) >= ma
    r_b = A.ElectronEventGenerator(fb, A.Material("Ar")).decays(365, 30.0)
    # sub-sample: rows 5000..5007 through the oracle with the replayed variates (row ids are global)
    sub = slice(5000, 5008)
    stream = []
    for r in range(sub.start, sub.stop):
        u = PX.uniform_1(11, PX.TAG_BREM_ROW, r, 1)[0]
        stream += [[u], PX.uniform_1(11, PX.TAG_BREM_EA, r, 1000)]
    oE, oF, info = O.brem_simulate(ef[sub], pfx, ma, g, O.Material("C"), 1000, O.UniformSource(np.concatenate(stream)))
    assert info["row_emit"].all()
    emit = fb._emit.cpu().numpy().astype(bool)
    first = int(emit[:sub.start].sum()) * 1000
    got_E = fb._E[first:first + 8000].cpu().numpy()
    got_F = fb._F[first:first + 8000].cpu().numpy()
    assert max_rel(got_E, oE) < 1e-12
    # the row weight interpolates Phi_e + Phi_p at E0, a table knot: identical whether the oracle sees 8 rows or 1e4
    assert max_rel(got_F, oF) < TOL
    # --- Resonance
    fr = A.FluxResonanceIsotropic(pfx, C_, axion_mass=ma, axion_coupling=g, n_samples=1_000_000, seed=12, **GEOM_DUNE)
    fr.simulate(); fr.propagate()
    u1, u2 = PX.uniform_2(12, PX.TAG_RESONANCE, 0, 1_000_000)
    _, oFr = O.resonance_simulate(pfx, ma, g, O.Material("C"), 1_000_000, O.UniformSource(np.concatenate([u1, u2])))
    assert max_rel(fr.axion_flux, oFr) < 1e-9
    # --- PairAnn
    fa = A.FluxPairAnnihilationIsotropic(pfx, C_, axion_mass=ma, axion_coupling=g, n_samples=100, seed=13, **GEOM_DUNE)
    fa.simulate(); fa.propagate(); fa.device_arrays()
    keep = O.pairann_row_valid(pfx, ma)
    assert fa._E.shape[0] == int(keep.sum()) * 100
    assert bool(torch.isfinite(fa._F).all()) and bool((fa._F >= 0).all())
    rows = np.nonzero(keep)[0][[0, 17, 4000, -1]]
    stream = []
    for r in rows:
        stream += [np.zeros(100), PX.uniform_1(13, PX.TAG_PAIRANN, int(r), 100)]
    oE, oF = O.pairann_simulate(pfx[rows], ma, g, O.Material("C"), 100, O.UniformSource(np.concatenate(stream)))
    idx = np.concatenate([np.arange(100) + 100 * int(np.searchsorted(np.nonzero(keep)[0], r)) for r in rows])
    assert max_rel(fa._E.cpu().numpy()[idx], oE) < TOL and max_rel(fa._F.cpu().numpy()[idx], oF) < TOL
    r_a = A.ElectronEventGenerator(fa, A.Material("Ar")).decays(365, 30.0)
    assert np.isfinite(r_b) and np.isfinite(r_a) and r_b >= 0 and r_a >= 0


def test_decay_4vectors_fp32_mode(A):
    """Optional single-precision decay MC: <= 1e-5 of the parent energy from the FP64 oracle on the same variates."""
    pf = np.stack([np.logspace(0.5, 4, 40), np.full(40, 1e12)], axis=1)
    ma, ns, seed = 0.1, 2003, 77
    fl = A.FluxPrimakoffIsotropic(pf, A.Material("W"), axion_mass=ma, axion_coupling=1e-5, seed=seed, **GEOM)
    fl.simulate(); fl.propagate()
    gen = A.PhotonEventGenerator(fl, A.Material("Ar"))
    l1, l2, w = gen.simulate_decay_4vectors(100, n_samples=ns, dtype=np.float32)
    assert l1.pmu.dtype == np.float32 and np.asarray(w).dtype == np.float32 and len(l1) == 40 * ns
    E = np.asarray(fl.axion_energy)
    us = [np.zeros(40)]
    for r in range(40):
        u1, u2 = PX.uniform_2_f32(seed, PX.TAG_DECAY2 + 16, r, ns)
        us += [u1, u2]
    o1, o2, ow = O.decay_4vectors(E, fl.decay_axion_weight, ma, 100, ns, O.UniformSource(np.concatenate(us)))
    scale = np.repeat(E, ns)[:, None]
    err = max(np.max(np.abs(l1.pmu - o1) / scale), np.max(np.abs(l2.pmu - o2) / scale))
    assert err < 1e-5, err
    assert np.max(np.abs(np.asarray(w) / ow - 1)) < 1e-6


def test_fp32_mode_weights(A, precision):
    """Optional single-precision mode: weights within 1e-5 of the FP64 oracle (north star) wherever the reference's own
    1 - exp(-x) is well conditioned (decay_prob > 1e-9; below that the reference loses digits, the expm1f path does not)."""
    if precision != "fast":
        pytest.skip("mode under test is selected explicitly")
    pf = _c2_flux(120)
    ma, g, ns, seed = 1.5, 3e-6, 400, 9
    fl = A.FluxComptonIsotropic(pf, A.Material("W"), axion_mass=ma, axion_coupling=g, n_samples=ns, seed=seed, precision="fp32",
                                keep_f64=True, **GEOM)
    fl.simulate(); fl.propagate()
    keep = O.compton_row_valid(pf, ma)
    u = np.concatenate([PX.uniform_1(seed, PX.TAG_COMPTON_EA, int(r), ns) for r in np.nonzero(keep)[0]])
    oE, oF = O.compton_simulate(pf, ma, g, 74, O.AbsCrossSection(O.Material("W")), ns, O.UniformSource(u))
    assert np.array_equal(fl.axion_energy, oE) and max_rel(fl.axion_flux, oF) < TOL          # production stays FP64
    p = O.propagate(oE, oF, ma, O.W_ee(g, ma), 1.0, 4.0, 0.2, geom_accept=GA)
    nz = p["scatter_w64"] != 0
    assert max_rel(fl.scatter_axion_weight_f64[nz], p["scatter_w64"][nz]) < 1e-5
    with np.errstate(all="ignore"):
        dec = np.where(nz, p["decay_w64"] / p["scatter_w64"], 0.0)
    ok = nz & (dec > 1e-9)
    assert ok.sum() > 1000
    assert max_rel(fl.decay_axion_weight_f64[ok], p["decay_w64"][ok]) < 1e-5
    r32 = A.ElectronEventGenerator(fl, A.Material("Ar")).decays(100, 5.0)
    _, r64 = O.decays(oE, p["decay_w"], 100, 5.0)
    assert r32 == pytest.approx(r64, rel=1e-5)
    # scan in the same mode
    from alplib_b200.scan import scan_rates
    gg = np.logspace(-7, -4, 5)
    m32 = scan_rates("compton", pf, A.Material("W"), [ma], gg, n_samples=ns, days_exposure=100, threshold=5.0, seed=seed, precision="fp32", **GEOM)
    m64 = scan_rates("compton", pf, A.Material("W"), [ma], gg, n_samples=ns, days_exposure=100, threshold=5.0, seed=seed, precision="fast", **GEOM)
    assert max_rel(m32, m64) < 1e-5


def test_examples_run(A, precision):
    """The example scripts (reference usage patterns with the import switched) run end to end."""
    if precision != "fast":
        pytest.skip("once is enough")
    import runpy
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runpy.run_path(os.path.join(root, "examples", "ex_readme_primakoff.py"), run_name="__main__")
    ns = runpy.run_path(os.path.join(root, "examples", "ex_na64_brem.py"), run_name="na64")
    energies, weights, n_decay, n_compton = ns["get_events"](10.0, 1e-6, use_loop=True)
    assert energies.shape == weights.shape == (10000,) and np.isfinite(n_decay) and np.isfinite(n_compton)
    assert n_decay >= 0 and n_compton >= 0 and np.sum(weights) == pytest.approx(n_decay + n_compton, rel=1e-6)
    ns = runpy.run_path(os.path.join(root, "examples", "ex_meson_decays.py"), run_name="__main__")
    assert len(ns["iso"].axion_energy) == 3 * 2 * 20000 and np.sum(ns["iso"].axion_flux) > 0
    assert 0 < len(ns["E_"]) <= 2000 * 2 * 500 and np.all(np.isfinite(ns["w"]))


# ---- charged-meson 3-body decay fluxes (fluxes.py:605-922) -----------------------------------------------------------------
MESON3_MODELS = ["scalar_ib1", "pseudoscalar_ib1", "vector_ib1", "scalar_ib2", "vector_ib2", "vector_ib9", "sd", "vector_contact"]


def test_meson3body_isotropic(A):
    """FluxChargedMeson3BodyIsotropic against the reference on the same variates: all interaction models, pion / kaon."""
    g = load_golden("meson3_iso")
    ns, c0, abd = int(g["n_samples"]), float(g["c0"]), tuple(float(x) for x in g["abd"])
    worst = 0.0
    for k in range(int(g["n_cases"])):
        mi, mt, ma, cpl = g[f"c{k}_cfg"]
        fl = A.FluxChargedMeson3BodyIsotropic(meson_flux=g["meson_flux"], boson_mass=float(ma), coupling=float(cpl),
                                              meson_type=["pion", "kaon"][int(mt)], interaction_model=MESON3_MODELS[int(mi)],
                                              n_samples=ns, c0=c0, abd=abd)
        fl.simulate(uniforms=g[f"c{k}_u"])
        assert max_rel(fl.axion_energy, g[f"c{k}_E"]) < 1e-13
        worst = max(worst, max_rel(fl.axion_flux, g[f"c{k}_F"]))
        fl.propagate()
        assert fl.scatter_axion_weight.dtype == np.float64 and np.all(fl.decay_axion_weight == 0)
        assert max_rel(fl.scatter_axion_weight, g[f"c{k}_S"]) < TOL
    assert worst < TOL, worst


def test_meson3body_decay_in_flight(A):
    """FluxChargedMeson3BodyDecay: decay positions, acceptance, energy cut, then propagate(decay=...) (f32 weights)."""
    g = load_golden("meson3_decay")
    ns = int(g["n_samples"])
    for k in range(int(g["n_cases"])):
        mi, mt, ma, cpl, cut = g[f"c{k}_cfg"]
        fl = A.FluxChargedMeson3BodyDecay(g["meson_flux"], boson_mass=float(ma), coupling=float(cpl), n_samples=ns,
                                          meson_type=["pion", "kaon"][int(mt)], interaction_model=MESON3_MODELS[int(mi)],
                                          energy_cut=float(g["energy_cut"]), keep_f64=True)
        fl.simulate(cut_on_solid_angle=bool(cut), uniforms=dict(pos=g[f"c{k}_upos"], samples=g[f"c{k}_usamp"]))
        assert max_rel(fl.axion_energy, g[f"c{k}_energy"]) < 1e-13
        assert max_rel(fl.axion_flux, g[f"c{k}_flux"]) < TOL
        assert max_rel(fl.cosines, g[f"c{k}_cosine"]) < 1e-12
        assert max_rel(fl.axion_angle, g[f"c{k}_angle"]) < TOL
        assert max_rel(fl.solid_angles, g[f"c{k}_solid_angle"]) < 1e-14
        assert max_rel(fl.decay_pos, g[f"c{k}_decay_pos"]) < 1e-14
        fl.propagate(decay="electron")
        assert fl.decay_axion_weight.dtype == np.float32
        assert max_rel(fl.scatter_axion_weight_f64, g[f"c{k}_s64"]) < TOL
        assert_decay_close(fl.decay_axion_weight_f64, g[f"c{k}_d64"], g[f"c{k}_s64"], TOL)
        assert max_rel(fl.scatter_axion_weight, g[f"c{k}_s32"]) <= F32
        fl.propagate()                                                    # no decays: float64 copies of the flux
        assert fl.scatter_axion_weight.dtype == np.float64 and max_rel(fl.scatter_axion_weight, g[f"c{k}_flux"]) < TOL


def test_meson3body_free_running(A):
    """Philox mode: reproducible, independent of row sharding, statistically consistent with the oracle's total rate."""
    mf = np.stack([np.linspace(50.0, 3000.0, 40), np.full(40, 0.5)], axis=1)
    kw = dict(boson_mass=5.0, coupling=1e-3, meson_type="kaon", interaction_model="scalar_ib1", n_samples=2000)
    a = A.FluxChargedMeson3BodyIsotropic(meson_flux=mf, seed=9, **kw); a.simulate()
    b = A.FluxChargedMeson3BodyIsotropic(meson_flux=mf, seed=9, **kw); b.simulate()
    assert np.array_equal(a.axion_flux, b.axion_flux) and np.array_equal(a.axion_energy, b.axion_energy)
    half = A.FluxChargedMeson3BodyIsotropic(meson_flux=mf[20:], seed=9, **kw)
    half.row_offset = 40                                                  # 20 mesons x 2 leptons precede this shard
    half.simulate()
    assert np.array_equal(half.axion_flux, a.axion_flux[40 * 2000:])
    us = O.UniformSource(np.random.RandomState(1))
    oE, oF = O.charged_meson_isotropic_simulate(mf[:4], 5.0, 1e-3, "kaon", "scalar_ib1", 2000, -0.97, [O.M_E, O.M_MU], (0, 0, 0), us)
    tot_gpu = np.sum(a.axion_flux[:4 * 2 * 2000])
    assert abs(tot_gpu / np.sum(oF) - 1) < 0.1                            # two independent MC estimates of the same rate
    # dGammadEa / total_br through the public methods
    d = a.dGammadEa(np.array([10.0, 100.0, 200.0]), O.M_MU)
    ref = [O.meson3_dgamma_dea(e, O.M_MU, 5.0, O.M_K, O.F_K, O.V_US, "scalar_ib1", -0.97, 1e-3) for e in (10.0, 100.0, 200.0)]
    assert max_rel(d, ref) < TOL
    assert a.total_br() > 0


def test_propagate_edge_energies(A):
    """The straight-line fast path must hand samples outside its domain to the general code: E_a == m_a (v = 0), E_a < m_a
    (NaN momentum), exponents far below -708, stable ALPs, huge energies -- element by element against the oracle, NaNs
    included, for odd and even lengths (scalar tail of the pair kernel)."""
    ma = 2.0
    e = np.array([ma, np.nextafter(ma, 3.0), ma * (1 + 1e-9), 0.5 * ma, 2.5, 10.0, 1e3, 1e8, 1e150, 3.0, 7.0])
    w = np.array([1.0, 2.0, 3.0, 4.0, 5e3, 6e-3, 7.0, 8.0, 9.0, 0.0, 1e300])
    for n in (e.shape[0], e.shape[0] - 1):
        for width in (0.0, 1e-20, 1e-12, 1e-6, 1.0):
            fl = A.FluxComptonIsotropic(np.array([[5.0, 1e9]]), A.Material("W"), axion_mass=ma, axion_coupling=1e-6,
                                        n_samples=2, keep_f64=True, **GEOM)
            fl.axion_energy = list(e[:n])
            fl.axion_flux = list(w[:n])
            A.AxionFlux.propagate(fl, width, 1.5)
            with np.errstate(all="ignore"):
                p = O.propagate(e[:n], w[:n], ma, width, 1.5, 4.0, 0.2)
            d, s_ = fl.decay_axion_weight_f64, fl.scatter_axion_weight_f64
            assert np.array_equal(np.isnan(s_), np.isnan(p["scatter_w64"])), (n, width, s_, p["scatter_w64"])
            assert np.array_equal(np.isnan(d), np.isnan(p["decay_w64"])), (n, width, d, p["decay_w64"])
            ok = ~np.isnan(p["scatter_w64"])
            assert max_rel(s_[ok], p["scatter_w64"][ok]) < TOL, (n, width)
            okd = ~np.isnan(p["decay_w64"])
            assert_decay_close(d[okd], p["decay_w64"][okd], p["scatter_w64"][okd], TOL)


def test_axion_meson_mixing(A):
    """FluxAxionMesonMixing: energy / angle filters on the device, both propagate branches (photon width; photon + hadronic)."""
    g = load_golden("meson_mixing")
    for k in range(int(g["n_cases"])):
        mt, mmass, ma, fa, new_fa = g[f"c{k}_cfg"]
        fl = A.FluxAxionMesonMixing(meson_flux=g[f"c{k}_mf"], meson_mass=float(mmass), axion_mass=float(ma), f_a=float(fa),
                                    meson_type=["Pi0", "Eta"][int(mt)], keep_f64=True)
        fl.simulate()
        assert np.array_equal(fl.axion_energy, g[f"c{k}_E"])
        assert max_rel(fl.axion_angle, g[f"c{k}_T"]) < 1e-13
        assert max_rel(fl.axion_flux, g[f"c{k}_F"]) < 1e-15
        assert fl.axion_coupling == pytest.approx(float(g[f"c{k}_g"][0]), rel=1e-15)
        fl.propagate()
        assert max_rel(fl.scatter_axion_weight_f64, g[f"c{k}_s64"]) < TOL
        assert_decay_close(fl.decay_axion_weight_f64, g[f"c{k}_d64"], g[f"c{k}_s64"], TOL)
        assert max_rel(fl.scatter_axion_weight, g[f"c{k}_s32"], floor=1e-44) <= F32
        fl.propagate(float(new_fa))
        assert max_rel(fl.scatter_axion_weight_f64, g[f"c{k}_s64n"]) < TOL
        assert_decay_close(fl.decay_axion_weight_f64, g[f"c{k}_d64n"], g[f"c{k}_s64n"], TOL)
        assert A.PhotonEventGenerator(fl, A.Material("Ar")).decays(100, 0.0) >= 0.0


@pytest.mark.parametrize("channel", ["brem", "brem_vector", "pairann"])
def test_epm_fused_equals_unfused(A, channel):
    """Call-by-call schedule vs the one-kernel chain for the e+- channels: identical bits for every per-sample array (the
    production arithmetic mode and the short-lived propagate variant are the same in every kernel of a launch)."""
    ef, pfx = _brem_tables(211)
    kw = dict(target_length=100, axion_mass=10.0, axion_coupling=1e-6, n_samples=500, seed=31, **GEOM_DUNE)

    def make(fused):
        if channel == "pairann":
            return A.FluxPairAnnihilationIsotropic(pfx, A.Material("C"), fused=fused, **kw)
        return A.FluxBremIsotropic(ef, pfx, A.Material("C"), fused=fused,
                                   boson_type="vector" if channel == "brem_vector" else "pseudoscalar", **kw)
    a = make(False)
    a.simulate(); a.propagate()
    ra = A.ElectronEventGenerator(a, A.Material("Ar")).decays(100, 5.0)
    b = make("auto")
    b.simulate(); b.propagate()
    rb = A.ElectronEventGenerator(b, A.Material("Ar")).decays(100, 5.0)
    assert len(a.axion_energy) > 1000
    for attr in ("axion_energy", "axion_flux", "decay_axion_weight", "scatter_axion_weight"):
        assert np.array_equal(getattr(a, attr), getattr(b, attr)), attr
    assert rb == pytest.approx(ra, rel=1e-12)
    c = make(True)
    c.simulate(); c.propagate()
    assert A.ElectronEventGenerator(c, A.Material("Ar")).decays(100, 5.0) == pytest.approx(ra, rel=1e-12)
    assert np.array_equal(c.axion_flux, a.axion_flux)              # regenerated on demand after the reduce-only chain


def test_flux_objects_are_freed_by_refcount(A):
    """No reference cycles in the flux / generator objects: dropping the last reference releases the (multi-GB at full
    size) device arrays at once instead of waiting for Python's cyclic garbage collector."""
    import gc
    import weakref
    gc.collect()
    gc.disable()
    try:
        for fused in ("auto", True, False):
            fl = A.FluxComptonIsotropic(_c2_flux(50), A.Material("W"), axion_mass=1.5, axion_coupling=1e-7, n_samples=200,
                                        seed=3, fused=fused, **GEOM)
            fl.simulate(); fl.propagate()
            gen = A.ElectronEventGenerator(fl, A.Material("Ar"))
            gen.decays(100, 5.0)
            r1, r2 = weakref.ref(fl), weakref.ref(gen)
            del fl, gen
            assert r1() is None and r2() is None, fused
        p = A.FluxPrimakoffIsotropic(_c2_flux(50), A.Material("W"), axion_mass=0.1, axion_coupling=1e-5, **GEOM)
        p.simulate()
        r = weakref.ref(p)
        del p
        assert r() is None
    finally:
        gc.enable()


def test_noncontiguous_flux_tables(A):
    """Strided / Fortran-ordered flux tables take the NumPy staging fallback (gather + packed pinned upload) instead of the
    native one-pass staging: identical samples either way, including tables with extra columns."""
    base = _c2_flux(120)
    wide = np.concatenate([base, np.arange(base.shape[0])[:, None] * 1.0], axis=1)        # an extra column, ignored
    variants = {"strided": np.repeat(base, 2, axis=0)[::2], "fortran": np.asfortranarray(base), "wide": wide,
                "wide_view": wide[:, :2]}
    kw = dict(axion_mass=1.5, axion_coupling=1e-7, n_samples=200, seed=5, **GEOM)
    ref = A.FluxComptonIsotropic(np.ascontiguousarray(base), A.Material("W"), **kw)
    ref.simulate()
    refp = A.FluxPrimakoffIsotropic(np.ascontiguousarray(base), A.Material("W"), axion_mass=1.5, axion_coupling=1e-5, **GEOM)
    refp.simulate()
    for name, tab in variants.items():
        c = A.FluxComptonIsotropic(tab, A.Material("W"), **kw)
        c.simulate()
        assert np.array_equal(c.axion_energy, ref.axion_energy) and np.array_equal(c.axion_flux, ref.axion_flux), name
        p = A.FluxPrimakoffIsotropic(tab, A.Material("W"), axion_mass=1.5, axion_coupling=1e-5, **GEOM)
        p.simulate()
        assert np.array_equal(p.axion_energy, refp.axion_energy) and np.array_equal(p.axion_flux, refp.axion_flux), name

"""CPU tier: the NumPy oracle against the golden fixtures produced by the unmodified reference
(tests/golden/make_golden.py) and the known-answer values of SURVEY.md section 8c."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, max_rel
from oracle import alp_oracle as O

KAT = json.load(open(os.path.join(GOLDEN, "kat.json")))
GEOM = dict(det_dist=4.0, det_length=0.2)
GA = O.geom_acceptance(0.04, 4.0)
GEOM_DUNE = dict(det_dist=574.0, det_length=5.0)
GA_DUNE = O.geom_acceptance(21.0, 574.0)


def test_constants_kat():
    # SURVEY section 8c "constants"
    assert O.METER_BY_MEV == 1.9733195759999998e-13 == KAT["METER_BY_MEV"]
    assert O.MEV2_CM2 == 3.8939901490248187e-22 == KAT["MEV2_CM2"]
    assert O.ALPHA == 0.0072992700729927005 == KAT["ALPHA"]
    assert O.HBARC == KAT["HBARC"] and O.AVOGADRO == KAT["AVOGADRO"] and O.M_E == KAT["M_E"]
    assert O.S_PER_DAY == KAT["S_PER_DAY"] and O.C_LIGHT == KAT["C_LIGHT"]


def test_scalar_kats():
    W = O.AbsCrossSection(O.Material("W"))
    C = O.AbsCrossSection(O.Material("C"))
    assert W.pe_data.shape[0] == 94 == KAT["absW.rows"] and C.pe_data.shape[0] == 80 == KAT["absC.rows"]
    pairs = [
        (O.primakoff_sigma(100, 1e-5, 0.1, 74), 2.2613107660719507e-08, "primakoff_sigma(100,1e-5,0.1,74)"),
        (W.sigma_mev(100), 0.06895240864110576, "absW.sigma_mev(100)"),
        (W.sigma_mev(1.0), 0.051874810225338346, "absW.sigma_mev(1.0)"),
        (W.sigma_mev(2e5), 2.5680599121454694e+21, "absW.sigma_mev(2e5)"),
        (O.W_gg(1e-5, 0.1), 4.973591971621732e-16, "W_gg(1e-5,0.1)"),
        (O.W_ee(1e-6, 10), 3.9580396838780663e-13, "W_ee(1e-6,10)"),
        (O.W_ee(1e-7, 1.5), 4.368643740120047e-16, "W_ee(1e-7,1.5)"),
        (O.W_gg_loop(1e-6, 10, O.M_E), 1.3054434139599113e-18, "W_gg_loop(1e-6,10,M_E)"),
        (O.W_gg_loop(1e-6, 0.5, O.M_E), 1.7700990709056778e-24, "W_gg_loop(1e-6,0.5,M_E)"),
        (O.compton_dsigma_dea(30, 100, 1e-5, 0.1, 74), 2.158825280625146e-15, "compton_dsigma_dea(30,100,1e-5,0.1,74)"),
        (O.compton_dsigma_dea(99.99, 100, 1e-5, 0.1, 74), 2.5868339784767592e-23, "compton_dsigma_dea(99.99,100,1e-5,0.1,74)"),
        (O.brem_dsigma_dea(300, 1000, 1e-6, 10, 6), 7.143121930424048e-22, "brem_dsigma_dea(300,1000,1e-6,10,6)"),
        (O.brem_dsigma_dx_vector(0.3, 1e-6, 10, 6), 1.0691941150993341e-17, "brem_dsigma_dx_vector(0.3,1e-6,10,6)"),
        (O.track_length_integrated_prob(1000, 400, 5), 0.001947957766158803, "track_length_integrated_prob(1000,400,5)"),
        (O.nu_integral(np.log(2.5), -1, 20 / 3), 2.5972770215450707, "nu_integral(ln2.5,-1,20/3)"),
        (O.icompton_sigma(np.array([50.0]), 0.1, 1e-5, 18)[0], 4.873988456052988e-13, "icompton_sigma([50],0.1,1e-5,18)"),
        (O.iprimakoff_sigma(100, 1e-5, 0.1, 18), 1.3379559091142171e-09, "iprimakoff_sigma(100,1e-5,0.1,18)"),
    ]
    for got, survey_value, key in pairs:
        assert KAT[key] == pytest.approx(survey_value, rel=1e-15), key       # fixture agrees with the survey probe
        assert float(got) == pytest.approx(survey_value, rel=2e-15), key     # oracle agrees with both
    assert O.W_ee(1e-7, 0.9) == 0.0 == KAT["W_ee(1e-7,0.9)"]


@pytest.mark.parametrize("mat", ["W", "C"])
def test_abs_table(mat):
    g = load_golden("abs_table_" + mat)
    a = O.AbsCrossSection(O.Material(mat))
    assert np.array_equal(a.pe_data[:, 0], g["table_e"]) and np.array_equal(a.pe_data[:, 1], g["table_xs"])
    assert max_rel(a.sigma_mev(g["energy"]), g["sigma_mev"]) == 0.0


def test_c1_primakoff_chain():
    g = load_golden("c1_primakoff")
    E, F = O.primakoff_simulate(g["photon_flux"], float(g["ma"]), float(g["g"]), 74, O.AbsCrossSection(O.Material("W")))
    assert F[0] == 327952.3965351194 == g["axion_flux"][0]
    p = O.propagate(E, F, 0.1, O.W_gg(1e-5, 0.1), 1.0, geom_accept=GA, **GEOM)
    assert p["decay_w"].dtype == np.float32
    assert p["decay_w"][0] == np.float32(3.2888147e-05) == g["decay_w32"][0]
    assert p["scatter_w"][0] == np.float32(65.2434) == g["scatter_w32"][0]
    ew, rate = O.decays(E, p["decay_w"], 100, 5.0)
    assert rate == 284.1535949707031 == float(g["decays_100_5"])
    assert max_rel(p["decay_w64"] * GA, g["decay_w64"]) < 1e-15
    sw, ipr = O.inverse_primakoff_events(E, p["scatter_w"], 1e-5, 0.1, 1e27, 0.04, 18, 100, 5.0)
    assert max_rel(ipr, g["iprimakoff_1e27_100_5"]) < 1e-15


def test_primakoff_ragged():
    g = load_golden("primakoff_ragged")
    ma, gg = float(g["ma"]), float(g["g"])
    E, F = O.primakoff_simulate(g["photon_flux"], ma, gg, 74, O.AbsCrossSection(O.Material("W")))
    assert np.array_equal(E, g["axion_energy"])
    assert max_rel(F, g["axion_flux"]) < 1e-15
    assert E.shape[0] < g["photon_flux"].shape[0]            # rows below ma were dropped
    p = O.propagate(E, F, ma, O.W_gg(gg, ma), 1.0, geom_accept=GA, **GEOM)
    assert max_rel(p["decay_w"], g["decay_w32"]) == 0.0 and max_rel(p["scatter_w"], g["scatter_w32"]) == 0.0
    nc = float(g["new_coupling"])
    p = O.propagate(E, F, ma, O.W_gg(nc, ma), np.power(nc / gg, 2), geom_accept=GA, **GEOM)
    assert max_rel(p["decay_w64"] * GA, g["decay_w64_new"]) < 1e-15
    assert max_rel(p["decay_w"], g["decay_w32_new"]) == 0.0


@pytest.mark.parametrize("name", ["katc_compton", "compton_mid", "compton_loop"])
def test_compton_chain(name):
    g = load_golden(name)
    ma, gg, ns = float(g["ma"]), float(g["g"]), int(g["n_samples"])
    us = O.UniformSource(g["uniforms"])
    E, F = O.compton_simulate(g["photon_flux"], ma, gg, 74, O.AbsCrossSection(O.Material("W")), ns, us)
    assert np.array_equal(E, g["axion_energy"])
    assert max_rel(F, g["axion_flux"]) < 1e-15
    loop = bool(g["loop_decay"]) if "loop_decay" in g else False
    iso = bool(g["is_isotropic"]) if "is_isotropic" in g else True
    width = O.W_ee(gg, ma) + (O.W_gg_loop(gg, ma, O.M_E) if loop else 0.0)
    p = O.propagate(E, F, ma, width, 1.0, geom_accept=GA if iso else None, **GEOM)
    assert max_rel(p["decay_w"], g["decay_w32"]) == 0.0 and max_rel(p["scatter_w"], g["scatter_w32"]) == 0.0
    assert max_rel(p["decay_w64"] * (GA if iso else 1.0), g["decay_w64"]) < 1e-15
    _, r = O.decays(E, p["decay_w"], 100, 5.0)
    assert max_rel(r, g["decays_100_5"]) < 1e-15
    _, r = O.compton_events(E, p["scatter_w"], gg, ma, 1e27, 0.04, 18, 100, 5.0)
    assert max_rel(r, g["compton_1e27_100_5"]) < 1e-15
    if "pair_weights" in g:
        pw, r = O.pair_production_events(E, p["scatter_w"], O.ALPPairProdutionCrossSection(O.Material("Ar"), ma), gg,
                                         1e27, 0.04, 100, 1.0)
        assert max_rel(pw, g["pair_weights"]) == 0.0 and r > 0


def test_katc_values():
    # SURVEY section 8c KAT-C literal values
    g = load_golden("katc_compton")
    assert list(g["axion_energy"]) == [14.38467993356044, 6.7935776965820205, 12.502295497863852, 28.238766300520226]
    np.testing.assert_allclose(g["axion_flux"], [9.9508379939949846e-08, 1.3118272264101634e-06,
                                                 1.1784504545853964e-06, 1.5785654301867164e-07], rtol=1e-15)
    assert float(g["decays_100_5"]) == pytest.approx(3.4814578064867874e-07, rel=1e-14)
    assert float(g["compton_1e27_100_5"]) == pytest.approx(6.752357257162864e-16, rel=1e-14)


def test_propagate_misc():
    g = load_golden("propagate_misc")
    E, F = g["axion_energy"], g["axion_flux"]
    p = O.propagate(E, F, 0.9, O.W_ee(1e-6, 0.9), 1.0, geom_accept=GA, **GEOM)
    assert np.all(p["decay_w"] == 0) and max_rel(p["decay_w"], g["zero_decay_w32"]) == 0.0
    assert max_rel(p["scatter_w"], g["zero_scatter_w32"]) == 0.0
    p = O.propagate(E, F, 0.9, float(g["tof_width"]), float(g["tof_rescale"]), timing_window=float(g["tof_window"]), **GEOM)
    assert max_rel(p["decay_w64"], g["tof_decay_w64"]) < 1e-15
    assert max_rel(p["scatter_w64"], g["tof_scatter_w64"]) < 1e-15
    assert (g["tof_decay_w64"] == 0).any() and (g["tof_decay_w64"] != 0).any()     # the cut bites on some samples


@pytest.mark.parametrize("name", ["katb_brem", "brem_ps", "brem_vec", "brem_notl"])
def test_brem(name):
    g = load_golden(name)
    us = O.UniformSource(g["uniforms"])
    boson = str(g["boson_type"]) if "boson_type" in g else "pseudoscalar"
    tl = bool(g["use_track_length"]) if "use_track_length" in g else True
    E, F, info = O.brem_simulate(g["electron_flux"], g["positron_flux"], float(g["ma"]), float(g["g"]), O.Material("C"),
                                 int(g["n_samples"]), us, boson_type=boson, use_track_length=tl)
    assert max_rel(E, g["axion_energy"]) < 1e-15
    assert max_rel(F, g["axion_flux"]) < 1e-13
    p = O.propagate(E, F, float(g["ma"]), O.W_ee(float(g["g"]), float(g["ma"])), 1.0, geom_accept=GA_DUNE, **GEOM_DUNE)
    assert max_rel(p["decay_w"], g["decay_w32"]) <= 1.2e-7
    assert max_rel(p["decay_w64"] * GA_DUNE, g["decay_w64"]) < 1e-13


def test_katb_values():
    g = load_golden("katb_brem")
    np.testing.assert_allclose(g["axion_energy"], [85.72610422711124, 16.654395595173387, 127.62316859164133,
                                                   242.77097612483072], rtol=1e-15)
    np.testing.assert_allclose(g["axion_flux"], [6.2151469459072448e-02, 3.8401145014387021e-05,
                                                 1.1293288193066638e-06, 1.4712097780461344e-05], rtol=1e-13)


def test_resonance():
    g = load_golden("katr_resonance")
    us = O.UniformSource(g["uniforms"])
    E, F = O.resonance_simulate(g["positron_flux"], 10.0, 1e-6, O.Material("C"), 1000, us)
    assert E[0] == 97.84735812133073 == g["axion_energy"][0]
    assert F[0] == pytest.approx(185.58954304351653, rel=1e-13)
    assert max_rel(F, g["axion_flux"]) < 1e-14


@pytest.mark.parametrize("name", ["kata_pairann", "pairann_mid"])
def test_pairann(name):
    g = load_golden(name)
    us = O.UniformSource(g["uniforms"])
    E, F = O.pairann_simulate(g["positron_flux"], 10.0, 1e-6, O.Material("C"), int(g["n_samples"]), us)
    assert max_rel(E, g["axion_energy"]) < 1e-14
    assert max_rel(F, g["axion_flux"]) < 1e-13
    if name == "kata_pairann":
        np.testing.assert_allclose(E, [206.6207601273814, 153.67083180362107, 890.0303251839387, 509.309777158082], rtol=1e-14)


@pytest.mark.parametrize("name", ["katd_decay4", "decay4_boosted", "decay4_heavy"])
def test_decay_4vectors(name):
    g = load_golden(name)
    us = O.UniformSource(g["uniforms"])
    p1, p2, w = O.decay_4vectors(g["axion_energy"], g["decay_w32"], float(g["ma"]), int(g["days"]), int(g["n_samples"]), us)
    assert np.array_equal(p1, g["p1"]) and np.array_equal(p2, g["p2"])          # bit-identical to the reference
    assert np.array_equal(w, g["w"])
    if name == "katd_decay4":
        assert tuple(p1[0]) == (46.15041488522627, -0.04616509485997197, 0.018813949523944135, 46.15038796042601)
        assert w[0] == 94.71786499023438 and w[3] == 26.58319664001465


def test_uniform_source_replays_legacy_numpy():
    us = O.UniformSource(np.random.RandomState(123))
    a = us.uniform(1.5, 20.0, 5)
    np.random.seed(123)
    b = np.random.uniform(1.5, 20.0, 5)
    assert np.array_equal(a, b)


def test_histogram_semantics():
    x = np.array([0.0, 0.5, 1.0, 1.0, 2.0, 2.5, -1.0, np.nan])
    w = np.arange(8, dtype=float) + 1
    h = O.histogram(x, w, [0.0, 1.0, 2.0])
    assert list(h) == [1 + 2, 3 + 4 + 5]          # last bin closed on the right, out-of-range and NaN dropped
    assert list(np.histogram(x, bins=[0.0, 1.0, 2.0], weights=w)[0]) == list(h)
    rs = np.random.RandomState(0)
    xr, wr = rs.uniform(-1, 11, 5000), rs.uniform(0.5, 1.5, 5000)
    for edges in (np.linspace(0, 10, 33), np.array([0.0, 0.3, 4.0, 4.5, 10.0])):
        np.testing.assert_allclose(O.histogram(xr, wr, edges), np.histogram(xr, bins=edges, weights=wr)[0], rtol=1e-12)


@pytest.mark.parametrize("name", ["volint_box", "volint_sphere", "volint_cyl"])
def test_propagate_iso_vol_int(name):
    g = load_golden(name)
    ma, gg, nc = float(g["ma"]), float(g["g"]), float(g["new_coupling"])
    geom = O.DetectorGeometry(float(g["l0"]), str(g["geometry"]), list(g["params"]), int(g["n_geom"]),
                              O.UniformSource(g["geom_uniforms"]))
    p = O.propagate_iso_vol_int(g["axion_energy"], g["axion_flux"], ma, O.W_ee(nc, ma), np.power(nc / gg, 2), 4.0, geom)
    assert max_rel(p["decay_w64"], g["decay_w64"]) < 1e-14 and max_rel(p["scatter_w64"], g["scatter_w64"]) < 1e-14
    assert max_rel(p["decay_w"], g["decay_w32"]) == 0.0


def test_nuclear_and_pi0():
    g = load_golden("nuclear")
    E, F = O.nuclear_simulate(g["rates"], float(g["ma"]), float(g["gann0"]), float(g["gann1"]), int(g["j"]), float(g["delta"]))
    assert np.array_equal(E, g["axion_energy"]) and max_rel(F, g["axion_flux"]) < 1e-15 and E.shape[0] == 3
    p = O.propagate(E, F, 0.3, O.W_ee(float(g["gae"]), 0.3), 1.0, geom_accept=GA, **GEOM)
    assert max_rel(p["scatter_w"], g["scatter_w32"]) == 0.0
    g = load_golden("pi0_flux")
    E, F = O.pi0_simulate_flux(g["momenta"], 30.0, 1e-3, 0.0259, O.UniformSource(g["uniforms"]), 50.0, 0.3)
    assert max_rel(E, g["axion_energy"]) < 1e-15 and max_rel(F, g["axion_flux"]) < 1e-15
    assert 0 < E.shape[0] < g["momenta"].shape[0]                  # the cuts bite


@pytest.mark.parametrize("name", ["meson2body", "meson2body_cut"])
def test_neutral_meson(name):
    g = load_golden(name)
    E, F, T = O.neutral_meson_simulate(g["meson_flux"], float(g["flux_weight"]), float(g["ma"]), float(g["g"]), O.M_PI0,
                                       int(g["n_samples"]), float(g["det_dist"]), float(g["det_area"]),
                                       O.UniformSource(g["uniforms"]), bool(g["angle_cut"]), float(g["off_axis"]))
    assert max_rel(E, g["axion_energy"]) < 1e-13 and max_rel(F, g["axion_flux"]) < 1e-14 and max_rel(T, g["axion_angle"]) < 1e-9
    if name == "meson2body_cut":
        assert 0 < E.shape[0] < g["meson_flux"].shape[0] * int(g["n_samples"])


MESON3_MODELS = ["scalar_ib1", "pseudoscalar_ib1", "vector_ib1", "scalar_ib2", "vector_ib2", "vector_ib9", "sd", "vector_contact"]


def test_meson3body_isotropic_oracle():
    """Charged-meson 3-body isotropic flux (fluxes.py:811-922): every interaction model, pion and kaon tables."""
    g = load_golden("meson3_iso")
    ns, c0, abd = int(g["n_samples"]), float(g["c0"]), tuple(g["abd"])
    for k in range(int(g["n_cases"])):
        mi, mt, ma, cpl = g[f"c{k}_cfg"]
        E, F = O.charged_meson_isotropic_simulate(g["meson_flux"], float(ma), float(cpl), ["pion", "kaon"][int(mt)],
                                                  MESON3_MODELS[int(mi)], ns, c0, [O.M_E, O.M_MU], abd,
                                                  O.UniformSource(g[f"c{k}_u"]))
        assert np.array_equal(E, g[f"c{k}_E"]) and np.array_equal(F, g[f"c{k}_F"])
        assert max_rel(F * (2 / (4 * np.pi * 20 ** 2)), g[f"c{k}_S"]) < 1e-15            # propagate: acceptance only (:917-920)
    assert any(np.any(g[f"c{k}_F"] != 0) for k in range(int(g["n_cases"])))


def test_meson3body_decay_oracle():
    """Decay-in-flight class (fluxes.py:605-808): exponential decay positions, solid-angle acceptance, energy cut."""
    g = load_golden("meson3_decay")
    ns = int(g["n_samples"])
    n_dead = 0
    for k in range(int(g["n_cases"])):
        mi, mt, ma, cpl, cut = g[f"c{k}_cfg"]
        # rebuild the flat stream in draw order from the structured fixture
        pos_u, blocks = g[f"c{k}_upos"], g[f"c{k}_usamp"].reshape(-1, 3 * ns)
        mm = O.MESON_PARAMS[["pion", "kaon"][int(mt)]][0]
        n_lep = sum(1 for ml in (O.M_E, O.M_MU) if not (ma > mm - ml))
        flat, b = [], 0
        for i in range(pos_u.shape[0]):
            flat.append(pos_u[i:i + 1])
            if g[f"c{k}_positions"][i] <= 52.0:
                flat.extend(blocks[b:b + n_lep]); b += n_lep
            else:
                n_dead += 1
        r = O.charged_meson_decay_simulate(g["meson_flux"], float(ma), float(cpl), ["pion", "kaon"][int(mt)],
                                           MESON3_MODELS[int(mi)], ns, float(g["c0"]), [O.M_E, O.M_MU],
                                           float(g["energy_cut"]), float(g["det_dist"]), float(g["det_length"]),
                                           O.UniformSource(np.concatenate(flat)), bool(cut))
        for key in ("energy", "cosine", "flux", "angle", "solid_angle", "decay_pos"):
            assert np.array_equal(r[key], g[f"c{k}_{key}"]), (k, key)
    assert n_dead > 0                                                  # some mesons decay beyond the dump


def test_meson_mixing_oracle():
    """FluxAxionMesonMixing (fluxes.py:1148-1240) restated: filters, constant rate, photon coupling, hadronic width."""
    g = load_golden("meson_mixing")
    for k in range(int(g["n_cases"])):
        mt, mmass, ma, fa, new_fa = g[f"c{k}_cfg"]
        E, T, F = O.meson_mixing_simulate(g[f"c{k}_mf"], float(ma), float(fa), float(mmass), ["Pi0", "Eta"][int(mt)], 1e21, 2.89,
                                          574.0, 21.0)
        assert np.array_equal(E, g[f"c{k}_E"]) and np.array_equal(T, g[f"c{k}_T"]) and np.array_equal(F, g[f"c{k}_F"])
        assert 0 < E.shape[0] < g[f"c{k}_mf"].shape[0]
        gc, wh = g[f"c{k}_g"]
        assert O.meson_mixing_photon_coupling(float(ma), float(fa)) == pytest.approx(gc, rel=1e-15)
        assert O.W_agg_hadronic(float(new_fa), float(ma)) == pytest.approx(wh, rel=1e-15, abs=0)
        p = O.propagate(E, F, float(ma), O.W_gg(gc, float(ma)), 1.0, 574.0, 5.0)
        assert max_rel(p["scatter_w64"], g[f"c{k}_s64"]) < 1e-14

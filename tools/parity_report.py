#!/usr/bin/env python
"""Measured parity of the CUDA path against the reference's golden fixtures, both arithmetic modes, written to
profiles/parity_r2.json (run on the GPU box)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import decay_tol, load_golden, max_rel   # noqa: E402
import alplib_b200 as A                                  # noqa: E402
from alplib_b200 import config                           # noqa: E402

GEOM = dict(det_dist=4.0, det_length=0.2, det_area=0.04)
DUNE = dict(det_dist=574.0, det_length=5.0, det_area=21.0)


def cond_err(d, d_ref, s_ref):
    """max over samples of |d - d_ref| / (|d_ref| * conditioned tolerance unit) -- in units of the 1e-9 + 4 ulp bound."""
    d, d_ref = np.asarray(d, float), np.asarray(d_ref, float)
    t = decay_tol(d_ref, s_ref, 1e-9, 4.0)
    nz = d_ref != 0
    if not nz.any():
        return 0.0
    return float(np.max(np.abs(d[nz] - d_ref[nz]) / (np.abs(d_ref[nz]) * t[nz])))


def report():
    out = {}
    for mode in ("faithful", "fast"):
        config.PRECISION = mode
        r = {}
        for name in ("c1_primakoff", "primakoff_ragged"):
            g = load_golden(name)
            fl = A.FluxPrimakoffIsotropic(g["photon_flux"], A.Material("W"), axion_mass=float(g["ma"]), axion_coupling=float(g["g"]), keep_f64=True, **GEOM)
            fl.simulate(); fl.propagate()
            ga = 0.04 / (4 * np.pi * 16.0)
            r[name] = {"flux": max_rel(fl.axion_flux, g["axion_flux"]), "scatter_w64": max_rel(fl.scatter_axion_weight_f64 * ga, g["scatter_w64"]),
                       "decay_w64_in_units_of_tolerance": cond_err(fl.decay_axion_weight_f64 * ga, g["decay_w64"], g["scatter_w64"]),
                       "scatter_w32": max_rel(fl.scatter_axion_weight, g["scatter_w32"])}
        for name in ("katc_compton", "compton_mid", "compton_loop"):
            g = load_golden(name)
            loop = bool(g["loop_decay"]) if "loop_decay" in g else False
            iso = bool(g["is_isotropic"]) if "is_isotropic" in g else True
            fl = A.FluxComptonIsotropic(g["photon_flux"], A.Material("W"), axion_mass=float(g["ma"]), axion_coupling=float(g["g"]),
                                        n_samples=int(g["n_samples"]), loop_decay=loop, is_isotropic=iso, keep_f64=True, **GEOM)
            fl.simulate(uniforms=g["uniforms"]); fl.propagate()
            ga = 0.04 / (4 * np.pi * 16.0) if iso else 1.0
            r[name] = {"energy_bit_exact": bool(np.array_equal(fl.axion_energy, g["axion_energy"])), "flux": max_rel(fl.axion_flux, g["axion_flux"]),
                       "scatter_w64": max_rel(fl.scatter_axion_weight_f64 * ga, g["scatter_w64"]),
                       "decay_w64_in_units_of_tolerance": cond_err(fl.decay_axion_weight_f64 * ga, g["decay_w64"], g["scatter_w64"]),
                       "scatter_w32": max_rel(fl.scatter_axion_weight, g["scatter_w32"]),
                       "scatter_w32_bit_identical_fraction": float(np.mean(np.asarray(fl.scatter_axion_weight) == g["scatter_w32"]))}
        for name in ("katb_brem", "brem_ps", "brem_vec", "brem_notl"):
            g = load_golden(name)
            boson = str(g["boson_type"]) if "boson_type" in g else "pseudoscalar"
            tl = bool(g["use_track_length"]) if "use_track_length" in g else True
            fl = A.FluxBremIsotropic(g["electron_flux"], g["positron_flux"], A.Material("C"), target_length=100, axion_mass=float(g["ma"]),
                                     axion_coupling=float(g["g"]), n_samples=int(g["n_samples"]), boson_type=boson, keep_f64=True, **DUNE)
            fl.simulate(use_track_length=tl, uniforms=g["uniforms"]); fl.propagate()
            r[name] = {"energy": max_rel(fl.axion_energy, g["axion_energy"]), "flux": max_rel(fl.axion_flux, g["axion_flux"])}
        g = load_golden("katr_resonance")
        fl = A.FluxResonanceIsotropic(g["positron_flux"], A.Material("C"), axion_mass=10.0, axion_coupling=1e-6, n_samples=1000, **DUNE)
        fl.simulate(uniforms=g["uniforms"])
        r["katr_resonance"] = {"flux": max_rel(fl.axion_flux, g["axion_flux"])}
        for name in ("kata_pairann", "pairann_mid"):
            g = load_golden(name)
            fl = A.FluxPairAnnihilationIsotropic(g["positron_flux"], A.Material("C"), axion_mass=10.0, axion_coupling=1e-6,
                                                 n_samples=int(g["n_samples"]), **DUNE)
            fl.simulate(uniforms=g["uniforms"])
            r[name] = {"energy": max_rel(fl.axion_energy, g["axion_energy"]), "flux": max_rel(fl.axion_flux, g["axion_flux"])}
        for name in ("katd_decay4", "decay4_boosted", "decay4_heavy"):
            g = load_golden(name)
            ns = int(g["n_samples"])
            fl = A.FluxPrimakoffIsotropic(g["photon_flux"], A.Material("W"), axion_mass=float(g["ma"]), axion_coupling=float(g["g"]), **GEOM)
            fl.simulate(); fl.propagate(); fl.decay_axion_weight = g["decay_w32"]
            npar = len(fl.axion_energy)
            l1, l2, w = A.PhotonEventGenerator(fl, A.Material("Ar")).simulate_decay_4vectors(int(g["days"]), n_samples=ns,
                                                                                            uniforms=g["uniforms"][npar:].reshape(npar, 2, ns))
            scale = np.repeat(np.asarray(fl.axion_energy), ns)[:, None]
            r[name] = {"max_dp_over_E_parent": float(max(np.max(np.abs(l1.pmu - g["p1"]) / scale), np.max(np.abs(l2.pmu - g["p2"]) / scale))),
                       "weights_bit_exact": bool(np.array_equal(np.asarray(w), g["w"]))}
        for name in ("volint_box", "volint_sphere", "volint_cyl"):
            g = load_golden(name)
            fl = A.FluxComptonIsotropic(g["photon_flux"], A.Material("W"), axion_mass=float(g["ma"]), axion_coupling=float(g["g"]), n_samples=12, keep_f64=True, **GEOM)
            fl.axion_energy = g["axion_energy"]; fl.axion_flux = g["axion_flux"]
            np.random.seed(78)
            geom = A.DetectorGeometry(float(g["l0"]), str(g["geometry"]), list(g["params"]), n_samples=int(g["n_geom"]))
            fl.propagate_iso_vol_int(geom, new_coupling=float(g["new_coupling"]))
            r[name] = {"decay_w64": max_rel(fl.decay_axion_weight_f64, g["decay_w64"]), "scatter_w64": max_rel(fl.scatter_axion_weight_f64, g["scatter_w64"])}
        # charged-meson 3-body decays and axion-meson mixing
        models = ["scalar_ib1", "pseudoscalar_ib1", "vector_ib1", "scalar_ib2", "vector_ib2", "vector_ib9", "sd", "vector_contact"]
        g = load_golden("meson3_iso")
        we, wf = 0.0, 0.0
        for k in range(int(g["n_cases"])):
            mi, mt, ma, cpl = g[f"c{k}_cfg"]
            fl = A.FluxChargedMeson3BodyIsotropic(meson_flux=g["meson_flux"], boson_mass=float(ma), coupling=float(cpl),
                                                  meson_type=["pion", "kaon"][int(mt)], interaction_model=models[int(mi)],
                                                  n_samples=int(g["n_samples"]), c0=float(g["c0"]), abd=tuple(float(x) for x in g["abd"]))
            fl.simulate(uniforms=g[f"c{k}_u"])
            we = max(we, max_rel(fl.axion_energy, g[f"c{k}_E"])); wf = max(wf, max_rel(fl.axion_flux, g[f"c{k}_F"]))
        r["meson3_iso (32 cases: 8 models x pion/kaon x masses)"] = {"energy": we, "flux": wf}
        g = load_golden("meson3_decay")
        wd = {"energy": 0.0, "flux": 0.0, "cosine": 0.0, "angle": 0.0}
        for k in range(int(g["n_cases"])):
            mi, mt, ma, cpl, cut = g[f"c{k}_cfg"]
            fl = A.FluxChargedMeson3BodyDecay(g["meson_flux"], boson_mass=float(ma), coupling=float(cpl), n_samples=int(g["n_samples"]),
                                              meson_type=["pion", "kaon"][int(mt)], interaction_model=models[int(mi)],
                                              energy_cut=float(g["energy_cut"]))
            fl.simulate(cut_on_solid_angle=bool(cut), uniforms=dict(pos=g[f"c{k}_upos"], samples=g[f"c{k}_usamp"]))
            for key, attr in (("energy", "axion_energy"), ("flux", "axion_flux"), ("cosine", "cosines"), ("angle", "axion_angle")):
                wd[key] = max(wd[key], max_rel(getattr(fl, attr), g[f"c{k}_{key}"]))
        r["meson3_decay (5 cases)"] = wd
        g = load_golden("meson_mixing")
        wm = {"angle": 0.0, "flux": 0.0, "scatter_w64": 0.0}
        for k in range(int(g["n_cases"])):
            mt, mmass, ma, fa, new_fa = g[f"c{k}_cfg"]
            fl = A.FluxAxionMesonMixing(meson_flux=g[f"c{k}_mf"], meson_mass=float(mmass), axion_mass=float(ma), f_a=float(fa),
                                        meson_type=["Pi0", "Eta"][int(mt)], keep_f64=True)
            fl.simulate(); fl.propagate()
            wm["angle"] = max(wm["angle"], max_rel(fl.axion_angle, g[f"c{k}_T"])); wm["flux"] = max(wm["flux"], max_rel(fl.axion_flux, g[f"c{k}_F"]))
            wm["scatter_w64"] = max(wm["scatter_w64"], max_rel(fl.scatter_axion_weight_f64, g[f"c{k}_s64"]))
        r["meson_mixing (4 cases)"] = wm
        # generic 2 -> 2 scatter MC (one arithmetic mode: the reference's operation order)
        g = load_golden("scatter2to2")
        ws = {"weights": 0.0, "cosines_abs": 0.0, "e3_lab": 0.0, "p3_lab_over_E": 0.0, "p4_lab_over_E": 0.0}
        for k in range(int(g["n_cases"])):
            pre = f"c{k}_"
            if bool(g[pre + "empty"]):
                continue
            model, ma, mN, z = str(g[pre + "model"]), float(g[pre + "ma"]), float(g[pre + "mN"]), int(g[pre + "z"])
            me = {"inverse_primakoff": lambda: A.M2InversePrimakoff(ma=ma, mN=mN, z=z), "primakoff": lambda: A.M2Primakoff(ma=ma, mN=mN, z=z),
                  "inverse_compton": lambda: A.M2InverseCompton(ma=ma, z=z), "compton": lambda: A.M2Compton(ma=ma, z=z),
                  "associated_production": lambda: A.M2AssociatedProduction(ma)}[model]()
            mc = A.Scatter2to2MC(me, p1=A.LorentzVector(*g[pre + "p1"]), p2=A.LorentzVector(*g[pre + "p2"]), n_samples=int(g[pre + "n"]))
            mc.scatter_sim(log_sampling=bool(g[pre + "log"]), uniforms=g[pre + "uniforms"])
            cos, w = mc.get_cosine_lab_weights()
            scale = abs(g[pre + "p1"][0]) + abs(g[pre + "p2"][0])
            ws["weights"] = max(ws["weights"], max_rel(w, g[pre + "weights"]))
            ws["cosines_abs"] = max(ws["cosines_abs"], float(np.max(np.abs(cos - g[pre + "cosines"]))))
            ws["e3_lab"] = max(ws["e3_lab"], max_rel(mc.get_e3_lab_weights()[0], g[pre + "e3_lab"]))
            ws["p3_lab_over_E"] = max(ws["p3_lab_over_E"], float(np.max(np.abs(mc.p3_lab_4vectors.pmu - g[pre + "p3_lab"])) / scale))
            ws["p4_lab_over_E"] = max(ws["p4_lab_over_E"], float(np.max(np.abs(mc.p4_lab_4vectors.pmu - g[pre + "p4_lab"])) / scale))
        r["scatter2to2 (9 cases: README snippet, 5 matrix elements, log sampling, moving target)"] = ws
        out[mode] = r
    out["tolerance"] = {"fp64_relative": 1e-9, "float32_attributes": 1.2e-7,
                        "decay_weights": "1e-9 + 4*2^-53/decay_prob (reported in units of that bound; <= 1 passes)"}
    worst = {m: max(v for c in out[m].values() for k, v in c.items() if isinstance(v, float) and "fraction" not in k and "units" not in k and "w32" not in k) for m in ("faithful", "fast")}
    out["worst_fp64_relative_error"] = worst
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_r2.json"), "w"), indent=1)
    print(json.dumps(out["worst_fp64_relative_error"]))
    for m in ("faithful", "fast"):
        for c, v in out[m].items():
            print(m, c, {k: (f"{x:.2e}" if isinstance(x, float) else x) for k, x in v.items()})


if __name__ == "__main__":
    report()

"""Scatter2to2MC (cross_section_mc.py:148-275) with the 2 -> 2 matrix elements of matrix_element.py:629-734.

CPU tier: the oracle restatement and the kernels' arithmetic (alp_math4.cuh built for the host) against the fixture the
unmodified reference produced (tests/golden/scatter2to2.npz, incl. the README.md:144-158 snippet).  GPU tier: the drop-in
classes (CUDA kernel through the C ABI) against the same fixture on the same injected uniforms, and the free-running
Philox stream replayed through the oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import dptr, load_golden, max_rel
from oracle import alp_oracle as O
from oracle import philox as PX

TOL = 1e-9


def _cases():
    g = load_golden("scatter2to2")
    out = []
    for k in range(int(g["n_cases"])):
        pre = f"c{k}_"
        out.append({key[len(pre):]: g[key] for key in g if key.startswith(pre)})
    return g, out


def _mtrx2(A, c):
    model, ma, mN, z = str(c["model"]), float(c["ma"]), float(c["mN"]), int(c["z"])
    return {"inverse_primakoff": lambda: A.M2InversePrimakoff(ma=ma, mN=mN, z=z),
            "primakoff": lambda: A.M2Primakoff(ma=ma, mN=mN, z=z),
            "inverse_compton": lambda: A.M2InverseCompton(ma=ma, z=z),
            "compton": lambda: A.M2Compton(ma=ma, z=z),
            "associated_production": lambda: A.M2AssociatedProduction(ma)}[model]()


def _check(c, w, cos, e3, p3cm, p3lab, p4lab, tol=TOL):
    scale = abs(c["p1"][0]) + abs(c["p2"][0])
    assert max_rel(w, c["weights"]) < tol
    assert np.max(np.abs(np.asarray(cos) - c["cosines"])) < tol
    assert max_rel(e3, c["e3_lab"]) < tol
    for got, ref in ((p3cm, c["p3_cm"]), (p3lab, c["p3_lab"]), (p4lab, c["p4_lab"])):
        assert np.max(np.abs(np.asarray(got) - ref)) / scale < tol


# ---- CPU tier ---------------------------------------------------------------------------------------------------------
def test_oracle_matches_reference_fixture():
    g, cases = _cases()
    assert len(cases) == 10
    for c in cases:
        masses, m2 = O.matrix_element(str(c["model"]), float(c["ma"]), mN=float(c["mN"]), z=int(c["z"]))
        r = O.scatter2to2_sim(masses, m2, c["p1"], c["p2"], int(c["n"]), O.UniformSource(c["uniforms"]), log_sampling=bool(c["log"]))
        if bool(c["empty"]):
            assert r is None
            continue
        _check(c, r["weights"], r["cosines"], r["e3_lab"], r["p3_cm"], r["p3_lab"], r["p4_lab"], tol=1e-13)
        assert np.array_equal(r["weights"], c["weights"])                    # pinned bit for bit on the reference
        assert max_rel(r["e4_lab"], c["e4_lab"]) < 1e-13
        assert np.max(np.abs(r["v3_lab"] - c["v3_lab"])) < 1e-14 and np.max(np.abs(r["v4_lab"] - c["v4_lab"])) < 1e-12
    # matrix elements evaluated directly
    sg, tg, mN = g["m2_s"], g["m2_t"], float(cases[0]["mN"])
    assert np.array_equal(np.array([O.m2_inverse_compton(s, tg, 0.4, 18, 1e-3) for s in sg]), g["m2_inverse_compton"])
    assert np.array_equal(np.array([O.m2_compton(s, tg, 0.4, 18, 1e-3) for s in sg]), g["m2_compton"])
    assert np.array_equal(np.array([O.m2_inverse_primakoff(s + mN ** 2, tg, 0.4, mN, 18, 1e-3) for s in sg]), g["m2_inverse_primakoff"])
    assert np.array_equal(np.array([O.m2_primakoff(s + mN ** 2, tg, 0.4, mN, 18, 1e-3) for s in sg]), g["m2_primakoff"])
    assert max_rel(np.array([O.m2_associated(s + 1.5, tg, 0.4) * 1e-6 for s in sg]), g["m2_associated"]) < 1e-15
    assert np.array_equal(O.atomic_elastic_ff2(g["atomic_ff_q"], 18), g["atomic_ff"])


def test_kernel_arithmetic_on_host_matches_reference_fixture(hostcheck):
    """alp_math4.cuh (what the CUDA kernel evaluates) compiled for the host, fed by the drop-in class's own parameter block."""
    import alplib_b200 as A
    g, cases = _cases()
    for c in cases:
        n = int(c["n"])
        mc = A.Scatter2to2MC(_mtrx2(A, c), p1=A.LorentzVector(*c["p1"]), p2=A.LorentzVector(*c["p2"]), n_samples=n)
        par = mc._params(bool(c["log"]))
        if bool(c["empty"]):
            assert par is None
            continue
        t, w, cos = np.empty(n), np.empty(n), np.empty(n)
        p3cm, p3lab, p4lab = np.empty((4, n)), np.empty((4, n)), np.empty((4, n))
        hostcheck.hc_scatter2(C.byref(par), C.c_uint64(n), dptr(np.ascontiguousarray(c["uniforms"])), dptr(t), dptr(w),
                              dptr(p3cm), dptr(p3lab), dptr(p4lab), dptr(cos))
        _check(c, w, cos, p3lab[0], p3cm.T, p3lab.T, p4lab.T)
    # direct matrix-element evaluation incl. coupling_product
    sg, tg, mN = g["m2_s"], g["m2_t"], float(cases[0]["mN"])
    for name, me, shift in (("m2_inverse_compton", A.M2InverseCompton(0.4, 18), 0.0), ("m2_compton", A.M2Compton(0.4, 18), 0.0),
                            ("m2_inverse_primakoff", A.M2InversePrimakoff(0.4, mN, 18), mN ** 2),
                            ("m2_primakoff", A.M2Primakoff(0.4, mN, 18), mN ** 2),
                            ("m2_associated", A.M2AssociatedProduction(0.4), 1.5)):
        for i, s in enumerate(sg):
            m = me._m2_struct(s + shift, 1e-3)
            out = np.empty(tg.shape[0])
            hostcheck.hc_m2_eval(C.byref(m), dptr(np.full(tg.shape[0], s + shift)), dptr(np.ascontiguousarray(tg)),
                                 C.c_uint64(tg.shape[0]), dptr(out))
            assert max_rel(out, g[name][i]) < 1e-12, name


def test_user_defined_matrix_element_is_rejected_not_run_on_the_cpu():
    import alplib_b200 as A

    class Mine(A.MatrixElement2):
        def __call__(self, s, t):
            return 1.0
    mc = A.Scatter2to2MC(Mine(0.1, 1.0, 0.0, 1.0), p1=A.LorentzVector(10, 0, 0, 9.9), p2=A.LorentzVector(1, 0, 0, 0), n_samples=4)
    with pytest.raises(NotImplementedError):
        mc.scatter_sim()
    assert A.MatrixElement2(1, 2, 3, 4)(1.0, 2.0) == 0.0          # the reference's base class returns 0.0


# ---- GPU tier ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def A():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    import alplib_b200
    from alplib_b200 import _native
    _native.lib()
    return alplib_b200


@pytest.mark.gpu
def test_scatter_sim_injected_uniforms_match_reference(A):
    g, cases = _cases()
    for c in cases:
        n = int(c["n"])
        mc = A.Scatter2to2MC(mtrx2=_mtrx2(A, c), p1=A.LorentzVector(*c["p1"]), p2=A.LorentzVector(*c["p2"]), n_samples=n)
        mc.scatter_sim(log_sampling=bool(c["log"]), uniforms=c["uniforms"])
        if bool(c["empty"]):
            assert len(mc.p3_lab_4vectors) == 0 and mc.dsigma_dcos_cm_wgts.shape[0] == 0
            assert mc.get_cosine_lab_weights()[0].shape[0] == 0
            continue
        cos, w = mc.get_cosine_lab_weights()
        e3, w3 = mc.get_e3_lab_weights()
        e4, _ = mc.get_e4_lab_weights()
        assert w is mc.dsigma_dcos_cm_wgts and w3 is w and len(mc.p3_lab_4vectors) == n
        _check(c, w, cos, e3, mc.p3_cm_4vectors.pmu, mc.p3_lab_4vectors.pmu, mc.p4_lab_4vectors.pmu)
        assert max_rel(e4, c["e4_lab"]) < TOL
        assert np.max(np.abs(np.asarray(mc.p3_lab_3vectors) - c["v3_lab"])) < TOL
        assert np.max(np.abs(np.asarray(mc.p4_lab_3vectors) - c["v4_lab"])) < TOL
        assert np.max(np.abs(np.asarray(mc.p3_cm_3vectors) - c["v3_cm"])) < TOL
        # the list-like attributes behave like the reference's lists of value objects
        v = mc.p3_lab_4vectors[3]
        assert isinstance(v, A.LorentzVector) and v.energy() == pytest.approx(c["e3_lab"][3], rel=TOL)
        assert mc.p3_lab_4vectors[3].cosine() == pytest.approx(c["cosines"][3], abs=TOL)
        assert isinstance(mc.p4_lab_3vectors[0], A.Vector3)
        # dsigma_dt through the device evaluator, and its integral (scipy quad driver, device integrand)
        assert max_rel(mc.dsigma_dt(float(c["s"]), c["t_probe"]), c["dsigma_dt_probe"]) < TOL
        if np.isfinite(c["total_xs"]) and str(c["tag"]) in ("icompton", "compton", "readme_iprim"):
            assert mc.get_total_xs(float(c["s"])) == pytest.approx(float(c["total_xs"]), rel=1e-8)
    # matrix elements are callable like the reference's
    sg, tg, mN = g["m2_s"], g["m2_t"], float(cases[0]["mN"])
    for i, s in enumerate(sg):
        assert max_rel(A.M2InverseCompton(0.4, 18)(s, tg, 1e-3), g["m2_inverse_compton"][i]) < 1e-12
        assert max_rel(A.M2Compton(0.4, 18)(s, tg, 1e-3), g["m2_compton"][i]) < 1e-12
        assert max_rel(A.M2InversePrimakoff(0.4, mN, 18)(s + mN ** 2, tg, 1e-3), g["m2_inverse_primakoff"][i]) < 1e-12
        assert max_rel(A.M2Primakoff(0.4, mN, 18)(s + mN ** 2, tg, 1e-3), g["m2_primakoff"][i]) < 1e-12
        assert max_rel(A.M2AssociatedProduction(0.4)(s + 1.5, tg, 1e-3), g["m2_associated"][i]) < 1e-12
    assert max_rel(A.AtomicElasticFF(18)(g["atomic_ff_q"]), g["atomic_ff"]) < 1e-14
    assert float(A.M2InverseCompton(0.4, 18)(50.0, -0.3)) == pytest.approx(float(O.m2_inverse_compton(50.0, -0.3, 0.4, 18)), rel=1e-13)


@pytest.mark.gpu
def test_readme_snippet_free_running_matches_oracle(A):
    """README.md:144-158 as written (free-running RNG): the kernel's Philox variates replayed through the oracle."""
    test_ma = 0.1
    matAr = A.Material("Ar")
    test_Ea = 100.0
    n, seed = 20000, 991
    m2_invPrim = A.M2InversePrimakoff(ma=test_ma, mN=matAr.m[0], z=matAr.z[0])
    mc = A.Scatter2to2MC(mtrx2=m2_invPrim, p1=A.LorentzVector(test_Ea, 0.0, 0.0, np.sqrt(test_Ea**2 - test_ma**2)),
                         p2=A.LorentzVector(matAr.m[0], 0.0, 0.0, 0.0), n_samples=n, seed=seed)
    mc.scatter_sim()
    cosines, dsdcos = mc.get_cosine_lab_weights()
    e3, dsde = mc.get_e3_lab_weights()
    u1, u2 = PX.uniform_2(seed, PX.TAG_SCATTER2, 0, n)
    masses, m2 = O.matrix_element("inverse_primakoff", test_ma, mN=float(matAr.m[0]), z=int(matAr.z[0]))
    r = O.scatter2to2_sim(masses, m2, mc.lv_p1.pmu, mc.lv_p2.pmu, n, O.UniformSource(np.concatenate([u1, u2])))
    assert max_rel(dsdcos, r["weights"]) < TOL and max_rel(e3, r["e3_lab"]) < TOL
    assert np.max(np.abs(cosines - r["cosines"])) < TOL
    # a second call draws a different stream (call counter in the Philox key); same seed + counter reproduces
    mc.scatter_sim()
    assert not np.array_equal(mc.dsigma_dcos_cm_wgts, dsdcos)
    mc2 = A.Scatter2to2MC(mtrx2=m2_invPrim, p1=mc.lv_p1, p2=mc.lv_p2, n_samples=n, seed=seed)
    mc2.scatter_sim()
    assert np.array_equal(mc2.dsigma_dcos_cm_wgts, dsdcos)
    # total cross section from the MC weights vs a fine deterministic quadrature of the oracle's dsigma/dt
    s = r["s"]
    p1c = np.sqrt((np.power(s - masses[0] ** 2 - masses[1] ** 2, 2) - np.power(2 * masses[0] * masses[1], 2)) / (4 * s))
    p3c = np.sqrt((np.power(s - masses[2] ** 2 - masses[3] ** 2, 2) - np.power(2 * masses[2] * masses[3], 2)) / (4 * s))
    assert np.sum(dsdcos) > 0 and p1c > 0 and p3c > 0

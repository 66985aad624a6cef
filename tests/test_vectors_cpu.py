"""CPU tier: the host-side value types keep the reference's interface (cross_section_mc.py:14-143)."""
import numpy as np

from alplib_b200.vectors import LorentzVector, Vector3


def test_vector3():
    a, b = Vector3(1.0, 2.0, 2.0), Vector3(0.0, 3.0, 4.0)
    assert a.mag() == 3.0 and a.mag2() == 9.0 and a * b == 14.0
    assert (a + b).vec.tolist() == [1.0, 5.0, 6.0] and (a - b).vec.tolist() == [1.0, -1.0, -2.0]
    assert (-a).vec.tolist() == [-1.0, -2.0, -2.0] and (2 * a).vec.tolist() == [2.0, 4.0, 4.0]
    u = a.unit_vec()
    assert abs(u.mag() - 1.0) < 1e-15 and Vector3(0, 0, 0).unit_vec().mag() == 0.0
    assert b.theta() == np.arccos(4.0 / 5.0) and b.phi() == np.arctan2(3.0, 0.0)
    assert str(a) == "(1.0,2.0,2.0)"


def test_lorentz_vector():
    p = LorentzVector(5.0, 0.0, 3.0, 4.0)
    assert p.mass2() == 0.0 and p.mass() == 0.0 and p.energy() == 5.0 and p.momentum() == 5.0
    assert p.cosine() == 0.8 and p.theta() == np.arccos(0.8) and p.phi() == np.arctan2(3.0, 0.0)
    q = LorentzVector(13.0, 0.0, 0.0, 12.0)
    assert q.mass() == 5.0 and p * q == 5.0 * 13.0 - 4.0 * 12.0
    s = p + q
    assert s.pmu.tolist() == [18.0, 0.0, 3.0, 16.0] and (s - q).pmu.tolist() == p.pmu.tolist()
    assert q.get_3velocity().vec.tolist() == [0.0, 0.0, 12.0 / 13.0] and q.get_3momentum().mag() == 12.0
    q.set_p4(1.0, 0.0, 0.0, 0.0)
    assert q.mass() == 1.0 and q.momentum3.mag() == 0.0

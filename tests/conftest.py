import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def max_rel(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) over elements where a != b (NaN == NaN, inf == inf)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    both_nan = np.isnan(a) & np.isnan(b)
    same = (a == b) | both_nan
    if same.all():
        return 0.0
    d = np.abs(a[~same] - b[~same])
    s = np.maximum(np.abs(b[~same]), max(floor, 1e-300))
    return float(np.max(d / s))


@pytest.fixture(scope="session", params=["libm", "seed21"])
def hostcheck(request):
    """Host build (g++, -ffp-contract=off) of the kernels' per-sample arithmetic -- test infrastructure only.  Two flavours:
    "libm" evaluates the fast mode's reciprocal / reciprocal square root exactly (1/a, 1/sqrt(d)); "seed21" runs the DEVICE
    algorithm -- the Newton correction steps of rcp_fast / rsqrt_fast -- on a seed cut to 21 mantissa bits, a little worse
    than the hardware's MUFU.RCP64H / RSQ64H, so the 1e-9 bar of the fast mode is checked on the CPU for the code the GPU runs."""
    d = os.path.join(ROOT, "tests", "hostcheck")
    seed = request.param == "seed21"
    so = os.path.join(d, "libhostcheck_seed21.so" if seed else "libhostcheck.so")
    deps = [os.path.join(d, "hostcheck.cpp"), os.path.join(d, "hostcheck2.inc"),
            os.path.join(ROOT, "alplib_b200", "csrc", "alp_math.cuh"),
            os.path.join(ROOT, "alplib_b200", "csrc", "alp_math2.cuh"),
            os.path.join(ROOT, "alplib_b200", "csrc", "alp_math3.cuh"),
            os.path.join(ROOT, "alplib_b200", "csrc", "alp_math4.cuh"), os.path.join(ROOT, "include", "alpk.h")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared"]
                              + (["-DALP_EMULATE_APPROX_SEED"] if seed else []) + ["-o", so, os.path.join(d, "hostcheck.cpp")])
    lib = ctypes.CDLL(so)
    lib.flavour = request.param
    return lib


def dptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


ULP1 = 2.0 ** -53      # spacing of doubles just below 1.0


def decay_tol(d_ref, s_ref, tol=1e-9, k=4.0):
    """Per-sample relative tolerance for decay weights.

    The reference evaluates decay_prob = 1 - exp(-x) (fluxes.py:62).  For long-lived ALPs x << 1 and the subtraction
    cancels: exp(-x) is quantised in steps of ulp(1)/2 = 2**-53, so the REFERENCE's own decay_prob carries an absolute
    error of that size and any libm whose exp differs in the last bit (NumPy's SIMD exp vs glibc vs CUDA libdevice, all
    <= 1 ulp) moves it by 2**-53 / decay_prob relative.  Parity is therefore asserted to tol + k * 2**-53 / decay_prob;
    for decay_prob >~ 1e-6 this is the plain 1e-9 of the north star."""
    d_ref = np.asarray(d_ref, dtype=np.float64)
    s_ref = np.asarray(s_ref, dtype=np.float64)
    with np.errstate(all="ignore"):
        dec = np.where(s_ref != 0, np.abs(d_ref / s_ref), 1.0)
        t = tol + k * ULP1 / np.maximum(dec, 1e-300)
    return np.where(np.isfinite(t), t, np.inf)


def assert_decay_close(d, d_ref, s_ref, tol=1e-9, k=4.0, extra_rel=0.0):
    d = np.asarray(d, dtype=np.float64)
    d_ref = np.asarray(d_ref, dtype=np.float64)
    assert d.shape == d_ref.shape
    t = decay_tol(d_ref, s_ref, tol, k) + extra_rel
    # (absolute floor 1e-300: products that land in the float64 denormal range carry no relative accuracy)
    bad = ~(np.abs(d - d_ref) <= t * np.abs(d_ref) + 1e-300) & ~(np.isnan(d) & np.isnan(d_ref))
    # decay_prob that cancels to exactly 0 in one libm and to 2**-53 in the other
    bad &= ~((np.minimum(np.abs(d), np.abs(d_ref)) == 0) & (np.maximum(np.abs(d), np.abs(d_ref)) <= k * ULP1 * np.abs(s_ref) * 1.0000001))
    assert not bad.any(), (int(bad.sum()), d[bad][:4], d_ref[bad][:4], t[bad][:4])

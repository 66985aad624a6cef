"""CPU tier: limit-setting helpers (reference fit.py:12-121 semantics) and the rate-map variant."""
import numpy as np

from alplib_b200 import fit


def _events(g):            # toy signal: rises as g^4, then dies off when the ALP decays before the detector
    return np.array([1e24 * g ** 4 * np.exp(-(g / 3e-5) ** 2)])


def test_two_sided_grid_search_and_map_variant():
    grid = np.logspace(-8, -3, 60)
    lo, hi = fit.TwoSidedGridSearch(_events, np.zeros(1), grid)
    stats = np.array([_events(g).sum() for g in grid])
    assert lo == grid[np.argmax(stats > 2.71)] and hi == grid[len(grid) - 1 - np.argmax(stats[::-1] > 2.71)]
    assert lo < hi
    rates = np.stack([stats, np.zeros_like(stats), stats * 1e-30])
    l, u = fit.limits_from_rate_map(grid, rates)
    assert l[0] == lo and u[0] == hi and np.isnan(l[1]) and np.isnan(u[2])
    assert fit.TwoSidedGridSearch(lambda g: np.zeros(1), np.zeros(1), grid) == (None, None)
    lo2, hi2 = fit.TwoSidedGridSearch(_events, np.zeros(1), grid, delta_chi2=False)      # stop value 2.3
    assert lo2 <= lo and hi2 >= hi


def test_binary_search_and_clean():
    x = fit.binary_search(lambda t: t ** 2, 9.0, 0.0, 10.0, tolerance=1e-6)
    assert abs(x - 3.0) < 1e-4
    x = fit.binary_search(lambda t: 100 - t, 40.0, 0.0, 100.0, tolerance=1e-6, is_increasing=False)
    assert abs(x - 60.0) < 1e-3
    m, lim = fit.cleanLimitData(np.array([1.0, 2.0, 3.0]), np.array([1e-6, 2e-6, 5e-5]), np.array([1e-4, 1e-4, 1e-5]))
    assert m.tolist() == [0.0, 1.0, 2.0, 2.0, 1.0] and lim.tolist() == [1e-6, 1e-6, 2e-6, 1e-4, 1e-4]


def test_limit_helpers_match_the_reference_fixture():
    """tests/golden/fit.json: TwoSidedGridSearch / binary_search / cleanLimitData as run in the unmodified reference
    (fit.py:12-121) by tests/golden/make_golden.py::case_fit on the same fixed generators."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fit.json")))
    grid, bkg, obs = np.array(g["grid"]), np.array(g["background"]), np.array(g["observations"])

    def events(x):
        return np.array([1e24 * x ** 4 * np.exp(-(x / 3e-5) ** 2), 3e23 * x ** 4 * np.exp(-(x / 1e-5) ** 2)])
    kws = {"signal_only": dict(observations=np.zeros(2)), "signal_only_cl": dict(observations=np.zeros(2), delta_chi2=False),
           "with_background": dict(observations=obs, background=bkg),
           "with_background_cl": dict(observations=obs, background=bkg, delta_chi2=False, target_cl=0.95)}
    seen = set()
    for k, kw in kws.items():
        lo, hi = fit.TwoSidedGridSearch(events, param_grid=grid, **kw)
        assert [float(lo), float(hi)] == g["TwoSidedGridSearch"][k], k
        seen.add((float(lo), float(hi)))
    assert len(seen) >= 3                    # the stop values 2.71 / 2.3 / chi2.ppf really select different grid points
    assert fit.TwoSidedGridSearch(lambda x: np.zeros(2), np.zeros(2), grid) == (None, None)
    assert g["TwoSidedGridSearch"]["no_signal"] == [None, None]
    bs = {"increasing": (lambda t: t ** 2, 9.0, 0.0, 10.0, dict(tolerance=1e-6)),
          "decreasing": (lambda t: 100 - t, 40.0, 0.0, 100.0, dict(tolerance=1e-6, is_increasing=False)),
          "coarse": (lambda t: np.exp(t), 5.0, -3.0, 4.0, dict(tolerance=0.1)),
          "edge": (lambda t: t, 1e300, 1.0, 1.0 + 4e-16, dict(tolerance=1e-3))}
    for k, (fn, stop, lo, hi, kw) in bs.items():
        assert float(fit.binary_search(fn, stop, lo, hi, **kw)) == g["binary_search"][k], k
    c = g["cleanLimitData"]
    for tag, smooth in (("plain", False), ("smooth", True)):
        m, lim = fit.cleanLimitData(np.array(c["masses"]), np.array(c["lower"]), np.array(c["upper"]), apply_smoothing=smooth)
        assert m.tolist() == c[tag]["masses"] and lim.tolist() == c[tag]["limits"], tag

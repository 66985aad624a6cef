"""Limit-setting helpers on top of the event-rate kernels (reference: fit.py:12-121).

`TwoSidedGridSearch`, `binary_search` and `cleanLimitData` keep the reference's signatures and semantics (they are
scalar host logic around a user-supplied `generator(param)`).  `limits_from_rate_map` applies the same two-sided
search to a whole (m_a, g) rate map from `alplib_b200.scan.scan_rates` -- one kernel launch instead of one full
simulate/propagate/decays chain per grid point."""
import numpy as np
from numpy import power
from scipy.signal import savgol_filter
from scipy.stats import chi2, chisquare


def TwoSidedGridSearch(generator, observations, param_grid, background=None, target_cl=0.90,
                       ddof=0, verbose=False, delta_chi2=True):
    """Brute-force scan of `param_grid` from both ends for the first points whose statistic exceeds the stop value;
    returns (lower, upper), or (None, None) when neither end moved (fit.py:12-58)."""
    def statistic(param):
        if background is not None:
            return np.sum(power(generator(param) + background - observations, 2)/background)
        elif background is not None and delta_chi2 == False:    # noqa: E712  (unreachable, as in the reference)
            return chisquare(generator(param) + background, observations, ddof)[0]
        return np.sum(generator(param))

    stop_value = chi2.ppf(target_cl, observations.shape[0]) - observations.shape[0] if background is not None else 2.3
    if delta_chi2:
        stop_value = 2.71
    lower_cl, upper_cl = _two_sided(param_grid, statistic, stop_value, verbose)
    if lower_cl == param_grid[0] and upper_cl == param_grid[-1]:
        return None, None
    return lower_cl, upper_cl


def _two_sided(param_grid, statistic, stop_value, verbose=False):
    lower_cl, upper_cl = param_grid[0], param_grid[-1]
    for x in param_grid:
        stat = statistic(x)
        if stat > stop_value:
            lower_cl = x
            if verbose:
                print("chi2 = ", stat, "found lower cl = ", lower_cl)
            break
    for x in param_grid[::-1]:
        stat = statistic(x)
        if stat > stop_value:
            upper_cl = x
            if verbose:
                print("chi2 = ", stat, "found upper cl = ", upper_cl)
            break
    return lower_cl, upper_cl


def limits_from_rate_map(g_grid, rates, stop_value=2.71):
    """Two-sided limits for every mass row of a rate map rates[n_ma, n_g] (total signal events per grid point):
    the TwoSidedGridSearch criterion `sum(generator(g)) > stop_value` with `generator` replaced by the precomputed map.
    Returns (lower[n_ma], upper[n_ma]) with NaN where the reference would return (None, None)."""
    g = np.asarray(g_grid, dtype=np.float64)
    r = np.asarray(rates, dtype=np.float64)
    above = r > stop_value
    any_above = above.any(axis=1)
    first = np.argmax(above, axis=1)
    last = r.shape[1] - 1 - np.argmax(above[:, ::-1], axis=1)
    lower = np.where(any_above, g[first], g[0])
    upper = np.where(any_above, g[last], g[-1])
    untouched = (lower == g[0]) & (upper == g[-1])
    return np.where(untouched, np.nan, lower), np.where(untouched, np.nan, upper)


def binary_search(test_function, stop_value, lower_edge, upper_edge, tolerance=0.1,
                  is_increasing=True, verbose=False):
    """Bisection for test_function(x) = stop_value on a monotonic test statistic (fit.py:63-100)."""
    x_upper, x_lower = upper_edge, lower_edge
    x = x_lower
    test_function(x)
    while abs(x_upper - x_lower) > tolerance*abs(upper_edge - lower_edge):
        test = test_function(x)
        if verbose:
            print("trying x, test_func = ", x, test)
        if (test < stop_value) == bool(is_increasing):
            x_lower = x
        else:
            x_upper = x
        x = (x_upper + x_lower)/2
        if x == x_lower or x == x_upper:
            if verbose:
                print("Ran into edge of search window, exiting")
            return x
    return x


def cleanLimitData(masses, lower_limits, upper_limits, apply_smoothing=False):
    """Join lower and upper limit curves into one closed contour (fit.py:105-121)."""
    bad = np.where(upper_limits - lower_limits < 0)
    upper_limits = np.delete(upper_limits, bad)
    lower_limits = np.delete(lower_limits, bad)
    masses = np.delete(masses, bad)
    if apply_smoothing:
        lower_limits = savgol_filter(lower_limits, 3, 0)
    joined_limits = np.append(lower_limits, upper_limits[::-1])
    joined_masses = np.append(masses, masses[::-1])
    joined_masses = np.append(0.0, joined_masses)
    joined_limits = np.append(joined_limits[0], joined_limits)
    return joined_masses, joined_limits

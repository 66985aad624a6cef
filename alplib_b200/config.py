"""Package-wide defaults."""

# "fast": algebraically simplified per-sample arithmetic, <= 1e-9 relative from the reference (the parity tolerance).
# "faithful": the reference's operation order expression by expression.
import os as _os

PRECISION = _os.environ.get("ALPK_PRECISION", "fast")          # (env override: benchmarking the other modes)

# Execution schedule of the drop-in classes (same results in every mode; the variates are a pure function of the seed):
#   "auto"  (default) simulate()/propagate() record their arguments; work is launched when a result is demanded.
#                     decays() on a Compton / Brem / pair-annihilation flux then runs ONE kernel that also materialises
#                     axion_energy, axion_flux, decay/scatter weights and event weights in HBM; reading an attribute
#                     earlier launches just what that attribute needs.
#   False             eager: every call launches its own kernel (the reference's call-by-call schedule).
#   True              deferred and reduce-only: decays() returns the rate without materialising per-sample arrays
#                     (they are regenerated bit-identically if read later).
FUSED = "auto"

# Strict compatibility with the reference's host containers (off by default: NumPy arrays are what vectorised user code
# wants).  True: `axion_energy` / `axion_flux` are Python lists of np.float64 (fluxes.py:222-224) and
# `simulate_decay_4vectors` returns three Python lists -- two of `LorentzVector` objects and one of weights
# (generators.py:124-128) -- so scripts that `.append`, concatenate with `+` or mutate the elements behave as with the
# reference.  Per object: the `strict=` keyword of the flux classes / `simulate_decay_4vectors(strict=...)`.
STRICT = _os.environ.get("ALPK_STRICT", "0") not in ("0", "", "false", "False")

# Rounds of the counter-based generator for the per-sample PRODUCTION streams (Compton / bremsstrahlung / associated-
# production E_a and angle draws, and the scan kernel that regenerates them).  10 = Philox4x32-10, the Random123 / cuRAND
# default; 7 = Philox4x32-7, the fewest rounds Salmon et al. (SC'11) found Crush-resistant.  The generator is a fifth of the
# chain kernel's issue slots, so the fast arithmetic mode defaults to 7 (6 % faster chain kernel); "faithful" keeps 10.  A
# variate stays a pure function of (seed, rounds, stream tag, row, sample); `oracle/philox.py` restates both.  Every other
# stream (decay angles, per-row draws, meson channels, 2 -> 2 scattering) is Philox4x32-10.
PHILOX_ROUNDS = {"fast": int(_os.environ.get("ALPK_PHILOX_ROUNDS", "7")), "fp32": int(_os.environ.get("ALPK_PHILOX_ROUNDS", "7")),
                 "faithful": 10}

# Device -> host copies of large per-sample arrays (flux.axion_energy ..., LorentzVectorArray.pmu): True returns NumPy arrays
# that are views of PINNED host memory from torch's caching host allocator, filled by one cudaMemcpyAsync at the PCIe rate
# (measured 52 GB/s on the B200 boxes); False copies through two 64 MiB pinned staging buffers into an ordinary (pageable)
# NumPy array, double-buffered -- bounded pinned memory, but limited by the host-side copy and the page faults of the fresh
# array (~5 GB/s measured).
PINNED_RESULTS = _os.environ.get("ALPK_PINNED_RESULTS", "1") not in ("0", "", "false", "False")

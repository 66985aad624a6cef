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

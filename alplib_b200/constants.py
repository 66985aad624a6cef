"""Physical constants and unit conversions in MeV / cm / s.

The numerical values are the reference's (constants.py:14-66, cited per line); they must agree bit for bit because
every weight is proportional to products of them (known-answer test in tests/).
Only the constants the Monte-Carlo flux-and-event path uses are defined.
"""
import numpy as np

# base units
GeV = 1.0e3
MeV = 1.0
keV = 1e-3

ALPHA = 1 / 137                        # constants.py:14 (fine-structure constant as used by the reference)
C_LIGHT = 2.99792458e10                # constants.py:20  cm/s
HBAR = 6.58212e-22                     # constants.py:21  MeV s
HBARC = 1.97236e-11                    # constants.py:22  MeV cm
AVOGADRO = 6.022e23                    # constants.py:29
M_E = 0.511                            # constants.py:32  MeV
M_P = 938.0                            # constants.py:33
M_MU = 105.658369                      # constants.py:34
M_PI = 139.57                          # constants.py:41
M_PI0 = 134.98                         # constants.py:42
M_ETA = 547.862                        # constants.py:43
M_ETA_PRIME = 957.78                   # constants.py:44
M_K = 493.677                          # constants.py:53  charged kaon
G_F = 1.16638e-11                      # constants.py:73  MeV^-2
CABIBBO = 0.9743                       # constants.py:74
V_UD = 0.97373                         # constants.py:78
V_US = 0.2243                          # constants.py:79
F_PI = 130.2                           # constants.py:91  pion decay constant
F_K = 155.7                            # constants.py:92  kaon decay constant
KAON_LIFETIME = 1.238e-8               # constants.py:105  s
PION_LIFETIME = 2.6e-8                 # constants.py:108  s
PION_WIDTH = 2.524e-14                 # constants.py:112  MeV
KAON_WIDTH = 5.317e-14                 # constants.py:114  MeV
METER_BY_MEV = 6.58212e-22 * 2.998e8   # constants.py:60  MeV m
MEV_PER_KG = 5.6095887e29              # constants.py:61
MEV2_CM2 = (METER_BY_MEV * 100) ** 2   # constants.py:65
S_PER_DAY = 3600 * 24                  # constants.py:66

pi = np.pi

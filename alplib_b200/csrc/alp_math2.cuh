// Per-sample arithmetic of the e+- channels (bremsstrahlung, resonance, associated production).
#pragma once
#include "alp_math.cuh"

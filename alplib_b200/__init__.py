"""alplib_b200 -- B200-native (sm_100a CUDA) implementation of alplib's Monte-Carlo flux-and-event pipeline.

Drop-in for the reference's hot path: `from alplib_b200 import *` gives FluxPrimakoffIsotropic, FluxComptonIsotropic,
FluxBremIsotropic, FluxResonanceIsotropic, FluxPairAnnihilationIsotropic, PhotonEventGenerator,
ElectronEventGenerator, Material, the decay widths and the constants -- same names and signatures as
`alplib.fluxes` / `alplib.generators`.  All per-sample work runs in hand-written CUDA kernels (libalpk.so, built by
`__graft_entry__.build()`); there is no CPU fallback.
"""
from .constants import *            # noqa: F401,F403
from .materials import Material, DetectorGeometry     # noqa: F401
from .decay import W_gg, W_ee, W_ff, W_gg_loop, W_agg_hadronic, fp2, b1   # noqa: F401
from .photon_xs import AbsCrossSection, ALPPairProdutionCrossSection   # noqa: F401
from .vectors import Vector3, LorentzVector, LorentzVectorArray   # noqa: F401
from .fluxes import *               # noqa: F401,F403
from .meson_fluxes import (FluxChargedMeson3BodyIsotropic, FluxChargedMeson3BodyDecay, M2Meson3BodyDecay,   # noqa: F401
                           FluxAxionMesonMixing)
from .generators import PhotonEventGenerator, ElectronEventGenerator   # noqa: F401
from .form_factors import AtomicElasticFF   # noqa: F401
from .matrix_element import (MatrixElement2, M2Compton, M2InverseCompton, M2InversePrimakoff, M2Primakoff,   # noqa: F401
                             M2AssociatedProduction)
from .cross_section_mc import Scatter2to2MC, lorentz_boost   # noqa: F401

__version__ = "0.1.0"

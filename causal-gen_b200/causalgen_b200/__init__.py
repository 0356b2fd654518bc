"""causalgen_b200: B200-native (sm_100a) drop-in for the HVAE image mechanism of biomedia-mira/causal-gen.

Public surface (mirrors reference src/vae.py, src/dmol.py, src/pgm/dscm.py):
    HVAE(args)            .forward / .abduct / .forward_latents / .sample
    DGaussNet, DmolNet    likelihood parameter heads
    counterfactual(...)   abduction -> action -> prediction combine of DSCM.forward
    CounterfactualGraph   the same, fixed batch, replayed from one CUDA graph
    vae_preprocess(...)   parent concatenation
"""
from .hvae import HVAE, CounterfactualGraph, counterfactual, ukbb_preprocess, vae_preprocess  # noqa: F401
from . import dp  # noqa: F401
from .model import DGaussNet, DmolNet  # noqa: F401

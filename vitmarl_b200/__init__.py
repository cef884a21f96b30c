"""vitmarl_b200 -- B200-native (sm_100a) rollout-and-encode hot path of hiepday3324/ViT-MARL.

Host-side mirror of the reference interfaces (``gymnax_exchange.jaxob.JaxOrderBookArrays``
function names, ``ExecutionAgent.normalize_vision_obs``, a flax-style ViT ``apply``) on top
of a C-ABI shared library of hand-written CUDA kernels.  There is no CPU fallback: every
op raises if the CUDA library is missing or no sm_100 GPU is present.
"""
from . import config  # noqa: F401

__version__ = "0.1.0"
__all__ = ["config", "jaxob", "vision", "synth", "vit", "env"]

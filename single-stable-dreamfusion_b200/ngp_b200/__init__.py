"""B200-native (sm_100a) NeRF rendering hot path of stable-dreamfusion.

``ngp_b200`` holds the CUDA kernels (``csrc/``), their C-ABI binding (``_cabi``) and the host-side
mirror of the reference's renderer / field network.  The sibling top-level modules
``gridencoder``, ``raymarching`` and ``freqencoder`` re-export the reference's exact Python
surface on top of it (drop-in: put this directory ahead of the reference on ``sys.path``).
"""
from . import _cabi  # noqa: F401

__all__ = ["_cabi"]

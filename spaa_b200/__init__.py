"""spaa_b200 -- B200-native (sm_100a) implementation of SPAA's hot path behind the reference's Python API.

Sub-modules mirror the reference's module names (/root/reference/src/python/): `models`, `pytorch_tps`,
`pytorch_ssim`, `perc_al` (+ `perc_al.differential_color_functions`), `projector_based_attack`, `train_network`,
`classifier`, `img_proc`.  All compute runs in hand-written CUDA kernels from libspaa_b200.so (include/spaa_b200.h);
importing this package loads that library and fails loudly if it has not been built.
"""
from ._lib import lib as _lib

_lib()          # dlopen libspaa_b200.so now: no silent fallback

from . import ops  # noqa: E402,F401

__all__ = ["ops"]

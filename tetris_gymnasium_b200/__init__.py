"""tetris_gymnasium_b200 -- B200-native batched Tetris simulator, drop-in for the hot path of
Max-We/Tetris-Gymnasium (env step/reset, 7-bag + queue + holder, grouped placements, observation wrappers).

The native library (libtetris_b200.so, hand-written sm_100a CUDA behind the C ABI in
include/tetris_b200.h) is loaded lazily by the env classes; importing this package works without a GPU.
"""
__version__ = "0.1.0"

from .mappings import ActionsMapping, RewardsMapping  # noqa: F401


def __getattr__(name):
    if name == "Tetris":
        from .envs.tetris import Tetris
        return Tetris
    raise AttributeError(name)

"""Constructor descriptors with the reference's component names (tetris_gymnasium/components/*).

On the B200 path the randomizer, queue and holder STATE lives in the per-env records in HBM and is advanced by the CUDA
kernels (csrc/tg_device.cuh: draw_piece / queue_pop / env_step); these classes only carry the constructor options, so that
`Tetris(randomizer=TrueRandomizer(7), queue=TetrominoQueue(r, size=7), holder=TetrominoHolder())` reads like the reference.
"""
from .tetromino import Pixel, Tetromino
from .tetromino_holder import TetrominoHolder
from .tetromino_queue import TetrominoQueue
from .tetromino_randomizer import BagRandomizer, Randomizer, TrueRandomizer

__all__ = ["Pixel", "Tetromino", "TetrominoHolder", "TetrominoQueue", "BagRandomizer", "Randomizer", "TrueRandomizer"]

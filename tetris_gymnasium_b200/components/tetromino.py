"""Pixel / Tetromino dataclasses -- mirrors tetris_gymnasium/components/tetromino.py:8-56 (constructor descriptors: the piece
geometry is compiled into the device tables by tg_create from `Tetris(tetrominoes=[...])`)."""
from dataclasses import dataclass

import numpy as np


@dataclass
class Pixel:
    id: int
    color_rgb: list


@dataclass
class Tetromino(Pixel):
    matrix: np.ndarray

    def __copy__(self):
        return Tetromino(id=self.id, color_rgb=list(self.color_rgb), matrix=np.array(self.matrix).copy())

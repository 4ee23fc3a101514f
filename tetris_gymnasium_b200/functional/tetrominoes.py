"""TETROMINOES constant table (reference functional/tetrominoes.py:42-147): ids 2..8, colours, and
matrices int8[7,4,4,4] = rot90(base, k=r) zero-padded bottom/right to 4x4."""
from dataclasses import dataclass

import numpy as np
import torch

_BASE = [
    (2, (0, 240, 240), [[0, 0, 0, 0], [1, 1, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]]),  # I
    (3, (240, 240, 0), [[1, 1], [1, 1]]),                                          # O
    (4, (160, 0, 240), [[0, 1, 0], [1, 1, 1], [0, 0, 0]]),                         # T
    (5, (0, 240, 0), [[0, 1, 1], [1, 1, 0], [0, 0, 0]]),                           # S
    (6, (240, 0, 0), [[1, 1, 0], [0, 1, 1], [0, 0, 0]]),                           # Z
    (7, (0, 0, 240), [[1, 0, 0], [1, 1, 1], [0, 0, 0]]),                           # J
    (8, (240, 160, 0), [[0, 0, 1], [1, 1, 1], [0, 0, 0]]),                         # L
]


@dataclass(frozen=True)
class Tetrominoes:
    base_pixels: torch.Tensor
    base_pixel_colors: torch.Tensor
    ids: torch.Tensor
    colors: torch.Tensor
    matrices: torch.Tensor


def _build():
    mats = np.zeros((7, 4, 4, 4), np.int8)
    for p, (_, _, m) in enumerate(_BASE):
        m = np.array(m, np.int8)
        for r in range(4):
            rm = np.rot90(m, k=r)
            mats[p, r, : rm.shape[0], : rm.shape[1]] = rm
    return Tetrominoes(
        base_pixels=torch.tensor([0, 1], dtype=torch.int8),
        base_pixel_colors=torch.tensor([[0, 0, 0], [128, 128, 128]], dtype=torch.uint8),
        ids=torch.tensor([t[0] for t in _BASE], dtype=torch.int8),
        colors=torch.tensor([t[1] for t in _BASE], dtype=torch.uint8),
        matrices=torch.from_numpy(mats),
    )


TETROMINOES = _build()


def get_tetromino_matrix(tetrominoes: Tetrominoes, tetromino_id: int, rotation: int) -> torch.Tensor:
    return tetrominoes.matrices[tetromino_id, rotation]

"""Functional API types -- mirrors tetris_gymnasium/functional/{core,tetrominoes}.py on torch tensors."""
from .core import EnvConfig, State  # noqa: F401
from .tetrominoes import TETROMINOES, Tetrominoes, get_tetromino_matrix  # noqa: F401

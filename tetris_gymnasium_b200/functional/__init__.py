"""Functional API types -- mirrors tetris_gymnasium/functional/{core,tetrominoes}.py on torch tensors."""
from .core import EnvConfig, State  # noqa: F401
from .tetrominoes import TETROMINOES, Tetrominoes, get_tetromino_matrix  # noqa: F401
from .queue import (bag_queue_get_next_element, create_bag_queue, create_uniform_queue,  # noqa: F401
                    uniform_queue_get_next_element)

"""CPU: the reference's base-env test suite (restated in tests/base_kats.py) on the C oracle.  The same cases run on the CUDA
env in tests/test_gpu_base_kats.py."""
import numpy as np
import pytest

import base_kats
from oracle.tetris_oracle import OracleEnv


class OracleAdapter(base_kats.Adapter):
    action_space_n = 8

    def __init__(self, gravity=True):
        self.o = OracleEnv(gravity=gravity)
        self.reset(seed=42)

    def _rot(self):
        p, m = self.o.scalars()["active"], self.o.active_matrix()
        return next(r for r in range(4) if np.array_equal(np.rot90(base_kats.BASE[p], k=r) > 0, m > 0))

    def reset(self, seed=42):
        return self.o.reset(seed=seed)[0]

    def step(self, a):
        return self.o.step(a)

    x = property(lambda s: s.o.scalars()["x"], lambda s, v: s.o.set_active(s.o.scalars()["active"], s._rot(), x=v))
    y = property(lambda s: s.o.scalars()["y"], lambda s, v: s.o.set_active(s.o.scalars()["active"], s._rot(), y=v))
    board = property(lambda s: s.o.board, lambda s, b: setattr(s.o, "board", b))
    game_over = property(lambda s: s.o.scalars()["game_over"])
    has_swapped = property(lambda s: s.o.scalars()["has_swapped"])

    def active_matrix(self):
        return self.o.active_matrix()

    def active_id(self):
        return self.o.scalars()["active"] + 2

    def set_active(self, piece, rot=0):
        self.o.set_active(piece, rot)

    def holder_ids(self):
        return [] if self.o.held_matrix() is None else [self.o.scalars()["holder"] + 2]


@pytest.mark.parametrize("case", base_kats.ALL, ids=lambda f: f.__name__)
def test_base_kat_on_oracle(case):
    if case is base_kats.kat_clone_restore_consistency:
        pytest.skip("the oracle has no whole-state snapshot (the CUDA env's get_state / set_state is what this case tests)")
    case(lambda gravity=True: OracleAdapter(gravity))

"""GPU: the reference's own base-env test suite (tests/test_base_env/**, restated in tests/base_kats.py) DIRECTLY on the CUDA env:
Tetris(num_envs=1, randomizer_mode="numpy") with the pokes of the reference tests (env.unwrapped.board / x / y / active_tetromino)
mapped onto tg_get_state / tg_set_state."""
import numpy as np
import pytest
import torch

import base_kats

pytestmark = pytest.mark.gpu


class GpuAdapter(base_kats.Adapter):
    def __init__(self, gravity=True):
        from tetris_gymnasium_b200.envs.tetris import Tetris

        self.env = Tetris(num_envs=1, gravity=gravity, randomizer_mode="numpy", autoreset_mode="disabled")
        self.action_space_n = self.env.action_space.n
        self.reset(seed=42)

    def _np(self, obs):
        return {k: v[0].cpu().numpy().copy() for k, v in obs.items()}

    def _s(self, key):
        return int(self.env.get_state()[key][0])

    def reset(self, seed=42):
        obs, _ = self.env.reset(seed=seed)
        return self._np(obs)

    def step(self, a):
        obs, r, term, trunc, info = self.env.step(torch.tensor([int(a)]))
        return self._np(obs), float(r[0]), bool(term[0]), bool(trunc[0]), {"lines_cleared": int(info["lines_cleared"][0])}

    x = property(lambda s: s._s("x"), lambda s, v: s.env.set_state(x=int(v)))
    y = property(lambda s: s._s("y"), lambda s, v: s.env.set_state(y=int(v)))
    board = property(lambda s: s.env.get_state()["board"][0].cpu().numpy().copy(), lambda s, b: s.env.set_state(board=np.asarray(b, np.uint8)))
    game_over = property(lambda s: bool(s._s("game_over")))
    has_swapped = property(lambda s: bool(s._s("has_swapped")))

    def active_matrix(self):
        st = self.env.get_state()
        p, r = int(st["piece"][0]), int(st["rotation"][0])
        return np.rot90(base_kats.BASE[p], k=r) * np.uint8(p + 2)

    def active_id(self):
        return self._s("piece") + 2

    def set_active(self, piece, rot=0):
        self.env.set_state(piece=int(piece), rotation=int(rot))

    def holder_ids(self):
        h = self._s("holder_piece")
        return [] if h < 0 else [h + 2]

    def snapshot(self):
        return self.env.get_state()

    def restore(self, snap):
        self.env.set_state(snap)

    def fingerprint(self):
        return tuple(t.cpu().numpy() for t in self.env.get_state()["_raw"])


@pytest.mark.parametrize("case", base_kats.ALL, ids=lambda f: f.__name__)
def test_base_kat_on_gpu(case):
    case(lambda gravity=True: GpuAdapter(gravity))


def test_seed_42_anchors_of_the_reference():
    """SURVEY appendix A, recorded from the live reference: reset(seed=42) -> S piece at x = 8, queue T L Z O."""
    e = GpuAdapter()
    obs = e.reset(seed=42)
    assert e.active_id() == 5 and (e.x, e.y) == (8, 0)
    assert obs["board"][0].tolist() == [1, 1, 1, 1, 0, 0, 0, 0, 0, 5, 5, 0, 0, 0, 1, 1, 1, 1]
    assert obs["active_tetromino_mask"].sum() == 9 and np.all(obs["holder"] == 1)
    assert obs["queue"][0].tolist() == [0, 4, 0, 0, 0, 0, 8, 0, 6, 6, 0, 0, 3, 3, 0, 0]

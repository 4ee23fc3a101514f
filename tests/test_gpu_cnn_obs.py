"""GPU parity of the fused CNN image adapter (tg_cnn_observe / CnnObservation) against the oracle chain
RgbObservation -> ResizeObservation(84, 84) -> GrayscaleObservation -> FrameStackObservation(4) (examples/train_cnn.py:127-147),
whose cv2 / gymnasium arithmetic is restated in oracle/cnn_obs_oracle.py and pinned against OpenCV."""
from collections import deque

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,shape", [
    (dict(width=10, height=20, gravity=True, queue_size=4), (84, 84)),
    (dict(width=20, height=40, gravity=True, queue_size=5), (84, 84)),     # BASELINE config 5
    (dict(width=10, height=20, gravity=True, queue_size=7), (84, 84)),
    (dict(width=24, height=28, gravity=True, queue_size=16), (84, 84)),    # image 32 x 96: x shrinks, y grows
    (dict(width=7, height=10, gravity=True, queue_size=2), (50, 37)),      # ragged sizes: plain-store path
], ids=lambda v: str(v) if isinstance(v, tuple) else f"{v['width']}x{v['height']}q{v['queue_size']}")
def test_cnn_observation_vs_oracle(cfg, shape):
    from gpu_util import OracleBatch, np_
    from oracle.cnn_obs_oracle import cnn_frame
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import CnnObservation

    n, T, L, K = 40, 60, 41, 4
    rng = np.random.default_rng(21)
    seqs = rng.integers(0, 7, size=(n, L)).astype(np.uint8)
    base = Tetris(num_envs=n, randomizer_mode="sequence", piece_sequences=seqs, autoreset_mode="next_step", **cfg)
    env = CnnObservation(base, shape=shape, stack_size=K, window=9, clip_reward=True)   # small window: exercises the slide
    orc = OracleBatch(n, seqs=seqs, **cfg)
    dsize = (shape[1], shape[0])

    def frames():
        return [cnn_frame(e.rgb(), dsize) for e in orc.envs]

    obs, _ = env.reset()
    orc.reset()
    stacks = [deque([f] * K, maxlen=K) for f in frames()]       # FrameStackObservation.reset: reset frame repeated
    assert tuple(obs.shape) == (n, K) + shape and obs.dtype == torch.uint8
    assert np.array_equal(np_(obs), np.stack([np.stack(s) for s in stacks]))
    n_reset = 0
    for t in range(T):
        a = rng.choice([0, 1, 2, 3, 5, 5, 6, 7], size=n)
        was_pending = orc.pending.copy()
        obs, r, term, trunc, info = env.step(torch.from_numpy(a))
        _, r2, t2, _ = orc.step(a)
        for i, f in enumerate(frames()):
            if was_pending[i]:          # NEXT_STEP autoreset: the wrapped env was reset -> the stack restarts
                stacks[i] = deque([f] * K, maxlen=K)
                n_reset += 1
            else:
                stacks[i].append(f)
        want = np.stack([np.stack(s) for s in stacks])
        got = np_(obs)
        if not np.array_equal(got, want):
            bad = np.argwhere((got != want).reshape(n, K, -1).any(2))
            raise AssertionError(f"t={t}: frame stacks differ at (env, slot) {bad[:6].tolist()}, max diff "
                                 f"{np.abs(got.astype(int) - want.astype(int)).max()}")
        assert np.array_equal(np_(r), np.sign(r2)) and np.array_equal(np_(term), t2)
    assert n_reset > 0
    base.close()

"""Property tests of the C oracle (CPU, hypothesis): the same size-independent invariants tests/test_gpu_fullsize.py checks
on the CUDA path at 4,194,304 envs, here on the checker itself over arbitrary action streams, board shapes and piece streams --
bedrock frame, no surviving full row, cells + W * lines == 0 (mod 4), observation = locked cells + the piece inside its mask
box, queue image = queue ids, feature identities of the grouped enumeration."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle.tetris_oracle import OracleEnv

SHAPES = [(10, 20, 4), (10, 20, 7), (20, 40, 5), (6, 8, 2), (13, 9, 3)]


@settings(max_examples=40, deadline=None)
@given(shape=st.sampled_from(SHAPES), gravity=st.booleans(), seed=st.integers(1, 2**31 - 1),
       actions=st.lists(st.integers(0, 7), min_size=20, max_size=400))
def test_base_env_invariants(shape, gravity, seed, actions):
    W, H, Q = shape
    env = OracleEnv(width=W, height=H, gravity=gravity, queue_size=Q)
    rng = np.random.default_rng(seed)
    env.set_sequence(rng.integers(0, 7, size=97))
    obs, _ = env.reset()
    lines_acc = 0
    for i, a in enumerate(actions):
        if i % 4 == 3:
            a = 5                                  # hard drops: commits and line clears
        obs, r, term, trunc, info = env.step(a)
        lines_acc += info["lines_cleared"]
        board = env.board
        field = board[:H, 4:4 + W]
        assert (board[H:, :] == 1).all() and (board[:H, :4] == 1).all() and (board[:H, 4 + W:] == 1).all()
        assert not (field != 0).all(axis=1).any()
        assert (int((field != 0).sum()) + W * lines_acc) % 4 == 0
        diff = obs["board"] != board
        assert int(diff.sum()) in (0, 4) and not (diff & (obs["active_tetromino_mask"] == 0)).any()
        s = env.scalars()
        n = 4 if s["active"] == 0 else (2 if s["active"] == 1 else 3)
        assert int(obs["active_tetromino_mask"].sum()) == n * n
        if diff.any():
            assert set(np.unique(obs["board"][diff])) == {s["active"] + 2}
        q = obs["queue"].reshape(4, Q, 4)
        assert (np.count_nonzero(q, axis=(0, 2)) == 4).all()
        assert not trunc and (r == 0.0 if term else True)
        if term:
            break


@settings(max_examples=25, deadline=None)
@given(shape=st.sampled_from([(10, 20, 4), (20, 40, 5), (7, 10, 4)]), seed=st.integers(1, 2**31 - 1), steps=st.integers(5, 80))
def test_grouped_feature_rows_are_self_consistent(shape, seed, steps):
    W, H, Q = shape
    env = OracleEnv(width=W, height=H, gravity=False, queue_size=Q)
    rng = np.random.default_rng(seed)
    env.set_sequence(rng.integers(0, 7, size=61))
    env.reset()
    for t in range(steps):
        feats, _, legal = env.grouped_observe(features=True, boards=False)
        h = feats[:, :W].astype(np.int64)
        assert (feats[:, W] == h.max(axis=1)).all()
        assert (feats[:, W + 2] == (np.abs(np.diff(h, axis=1)).sum(axis=1) & 255)).all()
        want = np.array([H - 1] * (W + 1) + [0, 0], np.uint8)
        assert (feats[legal == 0] == want).all() and legal.any()
        code, r, term, lines = env.grouped_step(int(rng.choice(np.flatnonzero(legal))), True)
        if term:
            break

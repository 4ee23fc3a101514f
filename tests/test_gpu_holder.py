"""GPU: TetrominoHolder(size > 1) (components/tetromino_holder.py:14-57) -- a FIFO of held pieces in the hot record -- against
the oracle (itself pinned against the live reference with a bigger holder assigned after construction,
oracle/validate_against_reference.py::check_holder): device step, host-buffer step (both modes), multi-step call, RGB image,
state round trip.  The "holder" observation has the fixed shape (P, P * size) where the reference's is ragged."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,cfg", [(2, dict()), (3, dict(gravity=False)), (4, dict(queue_size=7)), (2, dict(width=20, height=40, queue_size=5))])
def test_fifo_holder_vs_oracle(S, cfg):
    from tetris_gymnasium_b200.components import TetrominoHolder
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import RgbObservation
    from gpu_util import OracleBatch, assert_obs_equal, np_

    n = 96
    rng = np.random.default_rng(S)
    seqs = rng.integers(0, 7, size=(n, 128)).astype(np.uint8)
    env = Tetris(num_envs=n, holder=TetrominoHolder(S), randomizer_mode="sequence", piece_sequences=seqs, autoreset_mode="next_step", **cfg)
    host = Tetris(num_envs=n, holder=TetrominoHolder(S), randomizer_mode="sequence", piece_sequences=seqs, autoreset_mode="next_step", **cfg)
    rgbw = RgbObservation(env, keep_obs_dict=True)
    orc = OracleBatch(n, seqs=seqs, holder_size=S, **cfg)
    assert env.observation_space["holder"].shape == (4, 4 * S)
    obs, _ = env.reset()
    host.reset()
    assert_obs_equal(obs, orc.reset(), "reset")
    for t in range(150):
        a = rng.choice([0, 1, 2, 3, 5, 5, 6, 6, 6, 7], size=n)
        obs, r, term, _, info = env.step(torch.from_numpy(a))
        o2, r2, t2, l2 = orc.step(a)
        assert_obs_equal(obs, o2, f"t={t}")
        assert np.array_equal(np_(r), r2) and np.array_equal(np_(term), t2) and np.array_equal(np_(info["lines_cleared"]), l2)
        out = host.step_host(a.astype(np.int32), mode="compact" if t % 2 else "dma")
        for k in ("board", "active_tetromino_mask", "holder", "queue"):
            assert np.array_equal(out[k], o2[k]), (t, k)
        if t % 10 == 0:
            img = np_(rgbw.observation())
            for i in range(0, n, 7):
                assert np.array_equal(img[i], orc.envs[i].rgb()), (t, i)
    st = env.get_state()
    assert int(st["holder_count"].max()) <= S and st["holder_pieces"].shape == (n, S)
    twin = Tetris(num_envs=n, holder=TetrominoHolder(S), randomizer_mode="sequence", piece_sequences=seqs, autoreset_mode="next_step", **cfg)
    twin.reset()
    twin.set_state(st)
    a = torch.full((n,), 6, dtype=torch.int32, device="cuda")
    o1, _, _, _, _ = env.step(a)
    o2, _, _, _, _ = twin.step(a)
    for k in o1:
        assert torch.equal(o1[k], o2[k]), k
    # poking the FIFO through set_state
    twin.set_state(holder_pieces=torch.tensor([[2, 5] + [-1] * (S - 2)] * n), holder_rotations=torch.tensor([[1, 3] + [0] * (S - 2)] * n))
    st2 = twin.get_state()
    assert st2["holder_count"].tolist() == [2] * n and st2["holder_pieces"][0, :2].tolist() == [2, 5] and st2["holder_rotations"][0, :2].tolist() == [1, 3]


def test_fifo_holder_in_the_multi_step_call():
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from gpu_util import np_

    n, K = 3000, 24
    a, b = Tetris(num_envs=n, holder=3, queue_size=7), Tetris(num_envs=n, holder=3, queue_size=7)
    a.reset(seed=2); b.reset(seed=2)
    acts = torch.randint(5, 8, (K, n), dtype=torch.int32, device="cuda")
    obs, rew, term, _, _ = b.step_n(acts)
    for k in range(K):
        o1, r1, t1, _, _ = a.step(acts[k])
        assert torch.equal(o1["holder"], obs["holder"][k]) and torch.equal(o1["board"], obs["board"][k]) and torch.equal(r1, rew[k])

"""North-star acceptance: bit-exact trajectories on 10,000 SEEDED episodes (default 10x20 and the wide 20x40 board).

Every env i is seeded like the reference (`reset(seed=s+i)` -> PCG64(SeedSequence) 7-bag, numpy-exact on device) and
plays one full episode under a shared random action stream; after its game over it keeps being stepped exactly like
the unguarded reference.  The C oracle runs the same 10,000 envs (OpenMP) and every observation, reward,
termination flag and line count of every step is compared."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H,Q,p_drop,max_steps", [(10, 20, 4, 0.125, 700), (20, 40, 5, 0.45, 900)], ids=["default10x20", "wide20x40"])
def test_10k_seeded_episodes_bit_exact(W, H, Q, p_drop, max_steps):
    from gpu_util import np_
    from oracle.tetris_oracle import OracleVec
    from tetris_gymnasium_b200.envs.tetris import Tetris

    n, seed0 = 10_000, 20_000
    env = Tetris(width=W, height=H, queue_size=Q, num_envs=n, randomizer_mode="numpy", autoreset_mode="disabled")
    obs, _ = env.reset(seed=seed0)
    vec = OracleVec(n, width=W, height=H, gravity=True, queue_size=Q)
    for i, e in enumerate(vec.envs):
        e.reset(seed=seed0 + i)
    rng = np.random.default_rng(99)
    other = [a for a in range(8) if a != 5]
    done = np.zeros(n, bool)
    first = True
    total_lines = 0
    for t in range(max_steps):
        a = np.where(rng.random(n) < p_drop, 5, rng.choice(other, size=n)).astype(np.int32)
        obs, r, term, trunc, info = env.step(torch.from_numpy(a))
        vec.step(a, autoreset=False)
        if first or t % 1 == 0:
            assert np.array_equal(np_(obs["board"]), vec.board), t
            assert np.array_equal(np_(obs["active_tetromino_mask"]), vec.mask), t
            assert np.array_equal(np_(obs["holder"]), vec.holder), t
            assert np.array_equal(np_(obs["queue"]), vec.queue), t
        assert np.array_equal(np_(r), vec.reward), t
        assert np.array_equal(np_(term), vec.terminated.astype(bool)), t
        assert np.array_equal(np_(info["lines_cleared"]), vec.lines), t
        total_lines += int(vec.lines.sum())
        done |= vec.terminated.astype(bool)
        first = False
        if done.all():
            break
    assert done.all(), f"{(~done).sum()} of {n} episodes still running after {max_steps} steps"
    st = env.get_state()
    assert np.array_equal(np_(st["board"]), np.stack([e.board for e in vec.envs]))

"""GPU: device-side numpy seeding (tg_seed_numpy_seeds: SeedSequence hash + PCG64 srandom in CUDA) pinned against
np.random.PCG64(SeedSequence(s)).state for 10^5 seeds, and a 4 M-env reset(seed=...) in numpy mode in under a second."""
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_device_seeding_matches_numpy_for_1e5_seeds():
    from tetris_gymnasium_b200.envs.tetris import Tetris

    n = 100_000
    rng = np.random.default_rng(3)
    seeds = np.concatenate([np.arange(1, 40001, dtype=np.uint64), rng.integers(1, 1 << 32, 30000, dtype=np.uint64),
                            rng.integers(1 << 32, 1 << 63, 29996, dtype=np.uint64),
                            np.array([0xFFFFFFFF, 0x100000000, (1 << 64) - 1, 1 << 63], dtype=np.uint64)])
    env = Tetris(num_envs=n, randomizer_mode="numpy")
    env._seed_numpy(seeds)
    env._seeded = True
    torch.cuda.synchronize()
    rec = env._rng.cpu().numpy().reshape(n, env.layout.rng_stride)[:, :32].copy().view(np.uint64)   # state_hi, state_lo, inc_hi, inc_lo
    m64 = (1 << 64) - 1
    for i in rng.choice(n, 4000, replace=False).tolist() + [n - 4, n - 3, n - 2, n - 1]:
        st = np.random.PCG64(np.random.SeedSequence(int(seeds[i]))).state["state"]
        want = (st["state"] >> 64, st["state"] & m64, st["inc"] >> 64, st["inc"] & m64)
        assert tuple(int(v) for v in rec[i]) == want, int(seeds[i])
    # and in bulk against the vectorised restatement (all 10^5)
    from oracle.np_seed import pcg64_from_words, seed_words
    W = seed_words(seeds)
    for i in range(0, n, 997):
        assert tuple(int(v) for v in rec[i]) == pcg64_from_words(W[i])


def test_seeded_reset_of_4m_envs_is_fast_and_matches_the_oracle_stream():
    from oracle.tetris_oracle import OracleEnv
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from gpu_util import np_

    n = 1 << 22
    env = Tetris(num_envs=n, randomizer_mode="numpy", queue_size=7)
    env.reset(seed=1)                      # warm-up (allocations)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    obs, _ = env.reset(seed=12345)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert dt < 1.0, f"reset(seed) of {n} envs took {dt:.2f} s"
    q = np_(obs["queue"][:: n // 64][:64])
    b = np_(obs["board"][:: n // 64][:64])
    for k in range(64):
        i = k * (n // 64)
        o = OracleEnv(queue_size=7)
        want, _ = o.reset(seed=12345 + i)
        assert np.array_equal(q[k], want["queue"]) and np.array_equal(b[k], want["board"]), i

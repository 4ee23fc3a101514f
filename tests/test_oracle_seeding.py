"""CPU: the vectorised SeedSequence -> PCG64 restatement (oracle/np_seed.py, the arithmetic k_seed_numpy_seeds runs on the device)
against numpy itself, and the oracle's bulk seeding against its per-env seeding."""
import numpy as np

from oracle.np_seed import pcg64_from_words, seed_words
from oracle.tetris_oracle import OracleVec, numpy_pcg64_state


def test_seed_words_match_numpy_seedsequence():
    rng = np.random.default_rng(7)
    seeds = np.concatenate([np.arange(1, 30001, dtype=np.uint64), rng.integers(1, 1 << 32, 30000, dtype=np.uint64),
                            rng.integers(1 << 32, 1 << 63, 40000, dtype=np.uint64),
                            np.array([0xFFFFFFFF, 0x100000000, (1 << 64) - 1, 1 << 63], dtype=np.uint64)])
    W = seed_words(seeds)                       # 10^5 seeds, one vectorised pass
    for i in rng.choice(len(seeds), 3000, replace=False).tolist() + list(range(len(seeds) - 4, len(seeds))):
        ss = np.random.SeedSequence(int(seeds[i]))
        assert np.array_equal(ss.generate_state(4, np.uint64), W[i]), int(seeds[i])
        assert tuple(int(v) for v in numpy_pcg64_state(int(seeds[i]))) == pcg64_from_words(W[i]), int(seeds[i])


def test_bulk_oracle_equals_per_env_oracle():
    n = 512
    a = OracleVec(n, width=10, height=20, gravity=True, queue_size=7)
    for i, e in enumerate(a.envs):
        e.seed_numpy(1 + i)
        e.reset()
    b = OracleVec(n, bulk=True, width=10, height=20, gravity=True, queue_size=7)
    b.seed_all(1 + np.arange(n))
    b.reset_all()
    rng = np.random.default_rng(1)
    for t in range(80):
        act = rng.integers(0, 8, n).astype(np.int32)
        a.step(act)
        b.step(act)
        for k in ("board", "mask", "holder", "queue", "reward", "terminated", "lines"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (t, k)

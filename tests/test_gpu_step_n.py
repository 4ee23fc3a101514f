"""GPU: tg_step_n (K steps per native call; records resident in shared memory for small batches) against K tg_step calls."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _twin(n, **kw):
    from tetris_gymnasium_b200.envs.tetris import Tetris
    a, b = Tetris(num_envs=n, **kw), Tetris(num_envs=n, **kw)
    a.reset(seed=9); b.reset(seed=9)
    return a, b


@pytest.mark.parametrize("cfg,n,K", [
    (dict(queue_size=7), 4096, 16), (dict(queue_size=7), 1000, 7), (dict(queue_size=7), 65536, 8), (dict(queue_size=7), 33, 40),
    (dict(queue_size=4, gravity=False), 20000, 12), (dict(width=20, height=40, queue_size=5), 9000, 6),
    (dict(width=7, height=9, queue_size=3), 5000, 30), (dict(queue_size=7, autoreset_mode="same_step"), 3000, 50),
    (dict(queue_size=7, randomizer_mode="numpy"), 2048, 24), (dict(queue_size=7), 300000, 3),
])
def test_step_n_equals_k_steps(cfg, n, K):
    from gpu_util import np_

    one, many = _twin(n, **cfg)
    g = torch.Generator(device="cuda")
    g.manual_seed(n + K)
    for rep in range(3):
        acts = torch.randint(0, 8, (K, n), dtype=torch.int32, device="cuda", generator=g)
        obs, rew, term, trunc, info = many.step_n(acts, keep_all=True)
        for k in range(K):
            o1, r1, t1, _, i1 = one.step(acts[k])
            for key in ("board", "active_tetromino_mask", "holder", "queue"):
                assert np.array_equal(np_(o1[key]), np_(obs[key][k])), (rep, k, key)
            assert np.array_equal(np_(r1), np_(rew[k])) and np.array_equal(np_(t1), np_(term[k])), (rep, k)
            assert np.array_equal(np_(i1["lines_cleared"]), np_(info["lines_cleared"][k])) and not np_(trunc[k]).any()
        s1, s2 = one.get_state(), many.get_state()
        for raw1, raw2 in zip(s1["_raw"], s2["_raw"]):
            assert torch.equal(raw1, raw2), rep
        e1, e2 = one.episode_stats(), many.episode_stats()
        assert all(float(e1[k]) == float(e2[k]) for k in ("episodes", "sum_length", "sum_lines"))


def test_step_n_keep_last_only():
    from gpu_util import np_

    one, many = _twin(5000, queue_size=7)
    acts = torch.randint(0, 8, (20, 5000), dtype=torch.int32, device="cuda")
    obs, rew, term, _, info = many.step_n(acts, keep_all=False)
    for k in range(20):
        o1, r1, t1, _, i1 = one.step(acts[k])
    for key in ("board", "active_tetromino_mask", "holder", "queue"):
        assert np.array_equal(np_(o1[key]), np_(obs[key])), key
    assert np.array_equal(np_(r1), np_(rew)) and np.array_equal(np_(t1), np_(term))

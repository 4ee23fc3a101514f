"""GPU: Tetris(tetrominoes=[...]) (envs/tetris.py:88-89, 117-132) -- custom four-cell piece sets with their own colours and bag
size -- against the oracle (pinned against the live reference on the same sets: oracle/validate_against_reference.py::
check_custom_set).  Base env with the numpy-exact bag / TrueRandomizer, RGB image, host-buffer step, grouped features, and two
envs with DIFFERENT sets stepped alternately on one device (the piece tables are per device: the library swaps them)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sets():
    from oracle.validate_against_reference import CUSTOM_SETS
    return CUSTOM_SETS


def _tets(name):
    from tetris_gymnasium_b200.components import Tetromino
    mats, cols = _sets()[name]
    return [Tetromino(i, list(c), np.array(m, dtype=np.uint8)) for i, (m, c) in enumerate(zip(mats, cols))]


def _oracles(name, n, seeds, true_random=False, **kw):
    from oracle.tetris_oracle import OracleEnv
    mats, cols = _sets()[name]
    out = []
    for i in range(n):
        o = OracleEnv(**kw)
        o.set_tetrominoes(mats, cols)
        if true_random:
            o.set_true_randomizer()
        o.seed_numpy(int(seeds[i]))
        out.append(o)
    return out


@pytest.mark.parametrize("name,true_random,cfg", [("iot", False, dict()), ("odd5", False, dict(queue_size=7)), ("odd5", True, dict(width=12, height=16)),
                                                  ("only_i", False, dict()), ("odd5", False, dict(gravity=False))])
def test_custom_set_base_env_vs_oracle(name, true_random, cfg):
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import RgbObservation
    from gpu_util import np_

    n = 64
    seeds = 1000 + np.arange(n)
    mk = lambda: Tetris(num_envs=n, tetrominoes=_tets(name), randomizer_mode="numpy", randomizer="true" if true_random else None,  # noqa: E731
                        autoreset_mode="disabled", **cfg)
    env, host = mk(), mk()
    rgbw = RgbObservation(env, keep_obs_dict=True)
    orc = _oracles(name, n, seeds, true_random, **cfg)
    obs, _ = env.reset(seed=seeds)
    host.reset(seed=seeds)
    want = [o.reset()[0] for o in orc]
    keys = ("board", "active_tetromino_mask", "holder", "queue")
    for k in keys:
        assert np.array_equal(np_(obs[k]), np.stack([w[k] for w in want])), ("reset", k)
    assert env.observation_space["board"].high == 2 + len(_tets(name))
    rng = np.random.default_rng(3)
    for t in range(200):
        a = rng.integers(0, 8, size=n)
        obs, r, term, _, info = env.step(torch.from_numpy(a))
        out = host.step_host(a.astype(np.int32), mode="compact")
        res = [o.step(int(a[i])) if not o.scalars()["game_over"] or True else None for i, o in enumerate(orc)]
        for k in keys:
            w = np.stack([x[0][k] for x in res])
            assert np.array_equal(np_(obs[k]), w), (t, k)
            assert np.array_equal(out[k], w), ("host", t, k)
        assert np.array_equal(np_(r), np.array([x[1] for x in res], np.float32)) and np.array_equal(np_(term), np.array([x[2] for x in res]))
        if t % 20 == 0:
            img = np_(rgbw.observation())
            for i in range(0, n, 9):
                assert np.array_equal(img[i], orc[i].rgb()), (t, i)


@pytest.mark.parametrize("name", ["odd5", "iot"])
def test_custom_set_grouped_features_vs_oracle(name):
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations
    from gpu_util import np_

    n = 48
    seeds = 50 + np.arange(n)
    base = Tetris(num_envs=n, tetrominoes=_tets(name), gravity=False, randomizer_mode="numpy", autoreset_mode="disabled")
    env = GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)])
    orc = _oracles(name, n, seeds, gravity=False)
    g, info = env.reset(seed=seeds)
    for o in orc:
        o.reset()
    rng = np.random.default_rng(5)
    done = np.zeros(n, bool)
    for t in range(60):
        legal = np_(info["action_mask"])
        for i, o in enumerate(orc):
            if done[i]:
                continue
            f, _, lg = o.grouped_observe()
            assert np.array_equal(np_(g)[i], f) and np.array_equal(legal[i], lg), (t, i)
        a = np.array([rng.choice(np.flatnonzero(legal[i])) if legal[i].any() else 0 for i in range(n)])
        g, r, term, _, info = env.step(torch.from_numpy(a))
        for i, o in enumerate(orc):
            if done[i]:
                continue
            code, rr, tt, ll = o.grouped_step(int(a[i]))
            assert np.float32(rr) == np_(r)[i] and bool(tt) == bool(np_(term)[i]), (t, i)
            done[i] |= tt


def test_two_piece_sets_on_one_device_and_rollout_runs():
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from gpu_util import np_

    n = 256
    std, cus = Tetris(num_envs=n, randomizer_mode="numpy"), Tetris(num_envs=n, tetrominoes=_tets("odd5"), randomizer_mode="numpy")
    std2, cus2 = Tetris(num_envs=n, randomizer_mode="numpy"), Tetris(num_envs=n, tetrominoes=_tets("odd5"), randomizer_mode="numpy")
    for e in (std, cus, std2, cus2):
        e.reset(seed=9)
    acts = torch.randint(0, 8, (40, n), dtype=torch.int32, device="cuda")
    for t in range(40):             # alternate the two sets ...
        a, b = std.step(acts[t])[0], cus.step(acts[t])[0]
        ab, bb = {k: v.clone() for k, v in a.items()}, {k: v.clone() for k, v in b.items()}
        if t == 39:
            last = (ab, bb)
    for t in range(40):             # ... and run each on its own: same trajectories
        std2.step(acts[t])
    for t in range(40):
        cus2.step(acts[t])
    for k in last[0]:
        assert torch.equal(last[0][k], std2._obs()[k]) and torch.equal(last[1][k], cus2._obs()[k]), k
    g = Tetris(num_envs=n, tetrominoes=_tets("iot"), gravity=False, queue_size=7)
    g.reset(seed=1)
    g.rollout((-51, 76, -36, -18), 64)
    st = g.episode_stats()
    assert float(st["sum_length"]) >= 0 and int(g.get_state()["piece"].max()) <= 2


def test_unsupported_sets_fail_loudly():
    from tetris_gymnasium_b200.components import Tetromino
    from tetris_gymnasium_b200.envs.tetris import Tetris

    small = [Tetromino(0, [1, 2, 3], np.array([[1, 1], [1, 1]], dtype=np.uint8))]             # padding would be 2
    five = [Tetromino(0, [1, 2, 3], np.array([[1, 1, 1, 1], [1, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]], dtype=np.uint8))]
    for bad in (small, five):
        with pytest.raises(Exception):
            Tetris(num_envs=4, tetrominoes=bad)

"""GPU parity of the base env (tg_reset / tg_step through the Python host) against
(a) the golden trajectories recorded from the unmodified reference and (b) the C oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _episodes(z):
    for i in range(int(z["meta"][5])):
        yield {k[len(f"e{i}_"):]: z[k] for k in z.files if k.startswith(f"e{i}_")}


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "base_*.npz"))), ids=os.path.basename)
def test_golden_base_trajectories(path):
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, RgbObservation
    from gpu_util import np_

    z = np.load(path)
    W, H, gravity, Q, injected, _ = (int(v) for v in z["meta"])
    for ep in _episodes(z):
        kw = dict(width=W, height=H, gravity=bool(gravity), queue_size=Q, num_envs=1, autoreset_mode="disabled")
        if injected == 1:
            env = Tetris(randomizer_mode="sequence", piece_sequences=ep["seq"][None, :], **kw)
            obs, _ = env.reset()
        else:   # seeded numpy streams: 0 = 7-bag, 2 = TrueRandomizer
            env = Tetris(randomizer_mode="numpy", randomizer="true" if injected == 2 else None, **kw)
            obs, _ = env.reset(seed=int(ep["seed"]))
        featw, rgbw = FeatureVectorObservation(env), RgbObservation(env, keep_obs_dict=True)
        rgb_at = {int(t): i for i, t in enumerate(ep["rgb_t"])} if "rgb_t" in ep else {}
        T = len(ep["actions"])
        for t in range(T + 1):
            if t > 0:
                obs, r, term, trunc, info = env.step(torch.tensor([int(ep["actions"][t - 1])]))
                assert np_(r)[0] == ep["reward"][t - 1], (path, t)
                assert bool(np_(term)[0]) == bool(ep["terminated"][t - 1]) and not bool(np_(trunc)[0])
                assert int(np_(info["lines_cleared"])[0]) == int(ep["lines"][t - 1])
            assert np.array_equal(np_(obs["board"])[0], ep["board"][t]), (path, t)
            assert np.array_equal(np_(obs["active_tetromino_mask"])[0], ep["mask"][t]), (path, t)
            assert np.array_equal(np_(obs["holder"])[0], ep["holder"][t]), (path, t)
            assert np.array_equal(np_(obs["queue"])[0], ep["queue"][t]), (path, t)
            st = env.get_state()
            assert np.array_equal(np_(st["board"])[0], ep["locked"][t]), (path, t)
            assert (int(st["x"][0]), int(st["y"][0])) == (int(ep["x"][t]), int(ep["y"][t]))
            if "feat" in ep:
                assert np.array_equal(np_(featw.observation())[0], ep["feat"][t]), (path, t)
            if t in rgb_at:
                assert np.array_equal(np_(rgbw.observation())[0], ep["rgb"][rgb_at[t]]), (path, t)
        env.close()


@pytest.mark.parametrize("cfg", [
    dict(width=10, height=20, gravity=True, queue_size=4),
    dict(width=10, height=20, gravity=True, queue_size=7),
    dict(width=10, height=20, gravity=False, queue_size=4),
    dict(width=20, height=40, gravity=True, queue_size=5),
    dict(width=13, height=9, gravity=True, queue_size=3),
    dict(width=6, height=28, gravity=True, queue_size=1),
    dict(width=24, height=60, gravity=True, queue_size=16),
], ids=lambda c: f"{c['width']}x{c['height']}q{c['queue_size']}g{int(c['gravity'])}")
@pytest.mark.parametrize("autoreset", ["next_step", "disabled", "same_step"])
def test_random_actions_vs_oracle(cfg, autoreset):
    """Injected piece streams + random actions (all 8, incl. swap) on a ragged batch (n not a tile multiple)."""
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from gpu_util import OracleBatch, assert_obs_equal, np_

    n, T, L = 203, 260, 97
    if cfg["width"] >= 20:
        T = 400
    rng = np.random.default_rng(7)
    seqs = rng.integers(0, 7, size=(n, L)).astype(np.uint8)
    env = Tetris(num_envs=n, randomizer_mode="sequence", piece_sequences=seqs, autoreset_mode=autoreset, **cfg)
    orc = OracleBatch(n, seqs=seqs, **cfg)
    assert_obs_equal(env.reset()[0], orc.reset(), "reset")
    n_term = 0
    for t in range(T):
        a = rng.integers(0, 8, size=n)
        obs, r, term, trunc, info = env.step(torch.from_numpy(a))
        o2, r2, t2, l2 = orc.step(a, autoreset)
        assert_obs_equal(obs, o2, f"t={t}")
        assert np.array_equal(np_(r), r2), t
        assert np.array_equal(np_(term), t2) and not np_(trunc).any()
        assert np.array_equal(np_(info["lines_cleared"]), l2)
        n_term += int(t2.sum())
    assert n_term > 0
    st = env.get_state()
    assert np.array_equal(np_(st["board"]), np.stack([e.board for e in orc.envs]))
    assert not (np_(st["board"]) == 15).any()  # occupancy bitboards and id plane agree


def test_seeded_numpy_bag_matches_reference_stream():
    """randomizer_mode='numpy': reset(seed=s) reproduces BagRandomizer (PCG64 + Generator.shuffle) exactly,
    including the seed-42 anchors of SURVEY appendix A."""
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from gpu_util import OracleBatch, assert_obs_equal, np_

    k = np.load(os.path.join(GOLDEN, "reference_kat.npz"))
    env = Tetris(num_envs=1, randomizer_mode="numpy", autoreset_mode="disabled")
    obs, _ = env.reset(seed=42)
    assert np.array_equal(np_(obs["board"])[0], k["seed42_board"]) and np.array_equal(np_(obs["queue"])[0], k["seed42_queue"])
    assert np.array_equal(np_(obs["holder"])[0], k["seed42_holder"]) and np.array_equal(np_(obs["active_tetromino_mask"])[0], k["seed42_mask"])
    # 300 envs with seeds 1000+i: hard-drop only so that many bags are consumed
    n = 300
    env = Tetris(num_envs=n, randomizer_mode="numpy", autoreset_mode="next_step")
    orc = OracleBatch(n, seeds=1000 + np.arange(n))
    assert_obs_equal(env.reset(seed=1000)[0], orc.reset(), "reset")
    rng = np.random.default_rng(3)
    for t in range(120):
        a = rng.choice([0, 1, 5, 5, 6], size=n)
        obs, r, term, _, info = env.step(torch.from_numpy(a))
        o2, r2, t2, l2 = orc.step(a)
        assert_obs_equal(obs, o2, f"t={t}")
        assert np.array_equal(np_(term), t2)


def test_host_buffer_step_matches_device_step():
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from gpu_util import np_

    n = 1000
    rng = np.random.default_rng(5)
    seqs = rng.integers(0, 7, size=(n, 64)).astype(np.uint8)
    e1 = Tetris(num_envs=n, randomizer_mode="sequence", piece_sequences=seqs)
    e2 = Tetris(num_envs=n, randomizer_mode="sequence", piece_sequences=seqs)
    e1.reset(); e2.reset()
    bufs = e2.alloc_host_buffers(pinned=True)
    for t in range(40):
        a = rng.integers(0, 8, size=n).astype(np.int32)
        obs, r, term, trunc, info = e1.step(torch.from_numpy(a))
        out = e2.step_host(a, bufs)
        for k in ("board", "active_tetromino_mask", "holder", "queue"):
            assert np.array_equal(np_(obs[k]), out[k]), (t, k)
        assert np.array_equal(np_(r), out["reward"]) and np.array_equal(np_(term), out["terminated"].astype(bool))
        assert np.array_equal(np_(info["lines_cleared"]), out["lines_cleared"])


def test_philox_bag_properties_and_replay():
    """Device-native 7-bag: every aligned block of 7 draws is a permutation; feeding the drawn
    stream to the oracle reproduces the trajectory; results do not depend on the shard split."""
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from gpu_util import OracleBatch, assert_obs_equal, np_

    n = 64
    kw = dict(width=24, height=60, gravity=False, autoreset_mode="disabled", queue_size=1)
    env = Tetris(num_envs=n, **kw)
    obs, _ = env.reset(seed=11)
    streams = [[int(v)] for v in np_(env.get_state()["piece"])]
    q0 = np_(env.get_state()["queue"])[:, 0]
    for i in range(n):
        streams[i].append(int(q0[i]))
    done = False
    for k in range(400):                       # piece k drifts (k*7 % 12) cells left or right, then hard-drops
        for _ in range((k * 7) % 12):
            env.step(torch.full((n,), k % 2))
        _, _, term, _, _ = env.step(torch.full((n,), 5))
        if np_(term).any():
            break
        q = np_(env.get_state()["queue"])[:, 0]
        for i in range(n):
            streams[i].append(int(q[i]))
    nb = len(streams[0]) // 7
    assert nb >= 6, nb
    s = np.array(streams)[:, : nb * 7].reshape(n, -1, 7)
    assert (np.sort(s, axis=2) == np.arange(7)).all()
    assert len({tuple(r) for r in np.array(streams)}) > n // 2  # streams differ between envs
    # sharding invariance: envs 32..63 as a second shard with env_id_offset=32 give the same pieces
    env2 = Tetris(num_envs=32, env_id_offset=32, **kw)
    env2.reset(seed=11)
    assert np.array_equal(np_(env2.get_state()["piece"]), np.array(streams)[32:, 0])
    assert np.array_equal(np_(env2.get_state()["queue"])[:, 0], np.array(streams)[32:, 1])
    # replay through the oracle with the drawn streams injected
    seqs = np.array(streams, np.uint8)
    env3 = Tetris(num_envs=n, autoreset_mode="disabled", queue_size=1)
    orc = OracleBatch(n, seqs=seqs, queue_size=1)
    assert_obs_equal(env3.reset(seed=11)[0], orc.reset(), "reset")
    rng = np.random.default_rng(1)
    for t in range(100):
        act = rng.integers(0, 8, size=n)
        obs, r, term, _, _ = env3.step(torch.from_numpy(act))
        o2, r2, t2, _ = orc.step(act, "disabled")
        assert_obs_equal(obs, o2, f"t={t}")


def test_custom_action_and_reward_mappings_vs_oracle():
    """ActionsMapping / RewardsMapping are honoured like the reference: permuted action ids, two names sharing
    an id (first match in the elif chain wins; the hard_drop id also skips gravity), fractional rewards."""
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.mappings import ActionsMapping, RewardsMapping
    from gpu_util import OracleBatch, assert_obs_equal, np_

    cases = [
        (dict(move_left=7, move_right=6, move_down=5, rotate_clockwise=4, rotate_counterclockwise=3, hard_drop=2, swap=1, no_op=0),
         dict(alife=0.001, clear_line=1, game_over=-2.5, invalid_action=-0.1)),
        (dict(move_left=0, move_right=1, move_down=2, rotate_clockwise=3, rotate_counterclockwise=4, hard_drop=0, swap=6, no_op=7),
         dict(alife=1, clear_line=1, game_over=0, invalid_action=-0.1)),
        (dict(move_left=3, move_right=3, move_down=2, rotate_clockwise=5, rotate_counterclockwise=4, hard_drop=1, swap=6, no_op=1),
         dict(alife=0.3, clear_line=1, game_over=7.25, invalid_action=-1)),
    ]
    n, T = 120, 220
    for amap, rmap in cases:
        rng = np.random.default_rng(11)
        seqs = rng.integers(0, 7, size=(n, 71)).astype(np.uint8)
        env = Tetris(num_envs=n, randomizer_mode="sequence", piece_sequences=seqs, autoreset_mode="next_step",
                     actions_mapping=ActionsMapping(**amap), rewards_mapping=RewardsMapping(**rmap))
        assert env.reward_range == (min(rmap.values()), max(rmap.values()))
        orc = OracleBatch(n, seqs=seqs, actions=amap, alife=rmap["alife"], game_over=rmap["game_over"], invalid_action=rmap["invalid_action"])
        assert_obs_equal(env.reset()[0], orc.reset(), "reset")
        for t in range(T):
            a = rng.integers(0, 8, size=n)
            obs, r, term, _, info = env.step(torch.from_numpy(a))
            o2, r2, t2, l2 = orc.step(a)
            assert_obs_equal(obs, o2, f"t={t}")
            assert np.array_equal(np_(r), r2), (t, np_(r)[:8], r2[:8])
            assert np.array_equal(np_(term), t2) and np.array_equal(np_(info["lines_cleared"]), l2)


def test_masked_reset_and_state_roundtrip():
    """options={'reset_mask': ...} resets only the selected envs; get_state/set_state round-trip restores a snapshot
    (Tetris.get_state/set_state, envs/tetris.py:681-708)."""
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from gpu_util import np_

    n = 300
    rng = np.random.default_rng(2)
    seqs = rng.integers(0, 7, size=(n, 64)).astype(np.uint8)
    env = Tetris(num_envs=n, randomizer_mode="sequence", piece_sequences=seqs, autoreset_mode="disabled")
    env.reset()
    for t in range(40):
        env.step(torch.from_numpy(rng.integers(0, 8, size=n)))
    snap = env.get_state()
    acts = [torch.from_numpy(rng.integers(0, 8, size=n)) for _ in range(25)]
    outs = []
    for a in acts:
        obs, r, term, _, _ = env.step(a)
        outs.append((np_(obs["board"]).copy(), np_(r).copy(), np_(term).copy()))
    env.set_state(snap)                       # restore and replay: identical trajectory
    for a, (b, r0, t0) in zip(acts, outs):
        obs, r, term, _, _ = env.step(a)
        assert np.array_equal(np_(obs["board"]), b) and np.array_equal(np_(r), r0) and np.array_equal(np_(term), t0)
    before = env.get_state()
    mask = torch.from_numpy(rng.random(n) < 0.3)
    env.reset(options={"reset_mask": mask})
    after = env.get_state()
    m = mask.numpy()
    empty = np_(after["board"])[m][:, :20, 4:14]
    assert (empty == 0).all() and (np_(after["y"])[m] == 0).all()
    assert np.array_equal(np_(after["board"])[~m], np_(before["board"])[~m])
    assert np.array_equal(np_(after["x"])[~m], np_(before["x"])[~m]) and np.array_equal(np_(after["queue"])[~m], np_(before["queue"])[~m])


@pytest.mark.parametrize("cfg", [
    dict(width=10, height=20, gravity=True, queue_size=4),    # 24 x 34 px: 8-pixel pair-table path + TMA store
    dict(width=10, height=20, gravity=True, queue_size=7),
    dict(width=20, height=40, gravity=True, queue_size=5),    # BASELINE config 5: two-row 20-wide expansion
    dict(width=16, height=17, gravity=True, queue_size=1),    # 21 x 28 px: 4-pixel path, direct stores
    dict(width=7, height=10, gravity=True, queue_size=2),     # odd pixel count: generic path
    dict(width=13, height=30, gravity=False, queue_size=3),
], ids=lambda c: f"{c['width']}x{c['height']}q{c['queue_size']}")
def test_rgb_and_feature_wrappers_batched_vs_oracle(cfg):
    """RgbObservation / FeatureVectorObservation on a ragged batch during random play (swap included, so the holder
    panel is exercised) against the oracle's per-env rendering."""
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, RgbObservation
    from gpu_util import OracleBatch, np_

    n, T, L = 301, 120, 53
    rng = np.random.default_rng(11)
    seqs = rng.integers(0, 7, size=(n, L)).astype(np.uint8)
    env = Tetris(num_envs=n, randomizer_mode="sequence", piece_sequences=seqs, autoreset_mode="next_step", **cfg)
    rgbw, featw = RgbObservation(env, keep_obs_dict=True), FeatureVectorObservation(env)
    orc = OracleBatch(n, seqs=seqs, **cfg)
    env.reset()
    o2 = orc.reset()
    for t in range(T):
        if t % 6 == 0 or t > T - 4:
            got = np_(rgbw.observation())
            want = np.stack([e.rgb() for e in orc.envs])
            assert got.shape == want.shape and got.dtype == np.uint8
            if not np.array_equal(got, want):
                bad = np.flatnonzero((got != want).reshape(n, -1).any(1))
                raise AssertionError(f"t={t}: rgb differs for envs {bad[:8]} (of {len(bad)})")
            f = np_(featw.observation())
            fw = np.stack([e.features({k: v[i] for k, v in o2.items()}) for i, e in enumerate(orc.envs)])
            assert np.array_equal(f, fw), t
        a = rng.choice([0, 1, 2, 3, 4, 5, 5, 6, 7], size=n)
        env.step(torch.from_numpy(a))
        o2, _, _, _ = orc.step(a)
    env.close()


def test_true_randomizer_numpy_exact_and_philox():
    """TrueRandomizer (components/tetromino_randomizer.py:105-136): randomizer_mode='numpy' reproduces
    `default_rng(seed).integers(0, 7)` draw for draw (vs the oracle, itself pinned against numpy and the live reference);
    the device-native Philox variant is uniform, replayable and independent of the shard layout."""
    from tetris_gymnasium_b200.components import TetrominoQueue, TrueRandomizer
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from gpu_util import OracleBatch, assert_obs_equal, np_

    n = 200
    r = TrueRandomizer(7)
    env = Tetris(num_envs=n, randomizer_mode="numpy", queue=TetrominoQueue(r, size=5), autoreset_mode="next_step")
    assert env.randomizer_kind == "true" and env.queue_size == 5
    orc = OracleBatch(n, queue_size=5)
    for i, e in enumerate(orc.envs):
        e.set_true_randomizer()
        e.seed_numpy(500 + i)
    assert_obs_equal(env.reset(seed=500)[0], orc.reset(), "reset")
    rng = np.random.default_rng(9)
    for t in range(150):
        a = rng.choice([0, 1, 3, 5, 5, 5, 6], size=n)
        obs, rew, term, _, info = env.step(torch.from_numpy(a))
        o2, r2, t2, l2 = orc.step(a)
        assert_obs_equal(obs, o2, f"t={t}")
        assert np.array_equal(np_(rew), r2) and np.array_equal(np_(term), t2)
    # Philox: uniform pieces (no bag structure), same streams for the same (seed, global env id)
    n = 4096
    e1 = Tetris(num_envs=n, randomizer="true", queue_size=16, autoreset_mode="disabled")
    e1.reset(seed=3)
    q1 = np_(e1.get_state()["queue"])
    counts = np.bincount(q1.ravel(), minlength=7)
    assert counts.min() > 0.8 * q1.size / 7 and counts.max() < 1.2 * q1.size / 7
    assert (np.sort(q1[:, :7], axis=1) != np.arange(7)).any(axis=1).mean() > 0.9   # not permutations of 0..6
    e2 = Tetris(num_envs=n // 2, randomizer="true", queue_size=16, autoreset_mode="disabled", env_id_offset=n // 2)
    e2.reset(seed=3)              # per-env seed = seed + global env id: same (seed, id) pairs as the second half of e1
    assert np.array_equal(np_(e2.get_state()["queue"]), q1[n // 2:])


def test_record_episode_statistics_vector_format():
    """RecordEpisodeStatistics in the SyncVectorEnv info format (examples/train_lin_grouped.py:148, :316-322): episode
    return / length reported at the terminal step with the `_episode` mask, NEXT_STEP autoreset step not counted; the
    per-env sums agree with the oracle's trajectories and with the on-device totals (tg_stats)."""
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import RecordEpisodeStatistics
    from gpu_util import OracleBatch, np_

    n, T = 150, 400
    rng = np.random.default_rng(4)
    seqs = rng.integers(0, 7, size=(n, 33)).astype(np.uint8)
    base = Tetris(num_envs=n, randomizer_mode="sequence", piece_sequences=seqs, autoreset_mode="next_step")
    env = RecordEpisodeStatistics(base)
    orc = OracleBatch(n, seqs=seqs)
    env.reset(); orc.reset()
    base.episode_stats(reset=True)
    ret, length = np.zeros(n, np.float32), np.zeros(n, np.int64)
    n_eps, sum_r, sum_l = 0, 0.0, 0
    for t in range(T):
        a = rng.choice([0, 1, 3, 5, 5, 6], size=n)
        pend = orc.pending.copy()
        _, r, term, trunc, info = env.step(torch.from_numpy(a))
        _, r2, t2, _ = orc.step(a)
        ret[pend] = 0; length[pend] = 0
        ret[~pend] += r2[~pend]; length[~pend] += 1
        assert np.array_equal(np_(info["_episode"]), t2)
        assert np.array_equal(np_(info["episode"]["r"]), np.where(t2, ret, 0).astype(np.float32))
        assert np.array_equal(np_(info["episode"]["l"]), np.where(t2, length, 0))
        n_eps += int(t2.sum()); sum_r += float(ret[t2].sum()); sum_l += int(length[t2].sum())
    st = base.episode_stats()
    assert n_eps > 20 and int(st["episodes"]) == n_eps and int(st["sum_length"]) == sum_l and abs(float(st["sum_return"]) - sum_r) < 1e-3 * max(1.0, sum_r)

"""CPU: the C oracle (oracle/tetris_oracle.c) against the committed golden fixtures that were
recorded from the UNMODIFIED reference (oracle/make_golden.py), and against the reference's own
known-answer vectors.  This is what pins the oracle."""
import glob
import os

import numpy as np
import pytest

from oracle.tetris_oracle import OracleEnv, numpy_pcg64_state

from conftest import GOLDEN


def _episodes(z):
    n = int(z["meta"][5])
    for i in range(n):
        yield {k[len(f"e{i}_"):]: z[k] for k in z.files if k.startswith(f"e{i}_")}


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "base_*.npz"))), ids=os.path.basename)
def test_base_trajectories(path):
    z = np.load(path)
    W, H, gravity, Q, injected, _ = (int(v) for v in z["meta"])
    for ep in _episodes(z):
        env = OracleEnv(width=W, height=H, gravity=bool(gravity), queue_size=Q)
        if injected == 1:
            env.set_sequence(ep["seq"])
            obs, _ = env.reset()
        else:
            if injected == 2:   # seeded TrueRandomizer
                env.set_true_randomizer()
            obs, _ = env.reset(seed=int(ep["seed"]))
        T = len(ep["actions"])
        rgb_at = {int(t): i for i, t in enumerate(ep["rgb_t"])} if "rgb_t" in ep else {}
        for t in range(T + 1):
            if t > 0:
                obs, r, term, trunc, info = env.step(int(ep["actions"][t - 1]))
                assert np.float32(r) == ep["reward"][t - 1]
                assert term == bool(ep["terminated"][t - 1]) and trunc is False
                assert info["lines_cleared"] == int(ep["lines"][t - 1])
            assert np.array_equal(obs["board"], ep["board"][t]), (path, t)
            assert np.array_equal(obs["active_tetromino_mask"], ep["mask"][t])
            assert np.array_equal(obs["holder"], ep["holder"][t])
            assert np.array_equal(obs["queue"], ep["queue"][t])
            assert np.array_equal(env.board, ep["locked"][t])
            s = env.scalars()
            assert (s["x"], s["y"]) == (int(ep["x"][t]), int(ep["y"][t]))
            if "feat" in ep:
                assert np.array_equal(env.features({k: v.copy() for k, v in obs.items()}), ep["feat"][t])
            if t in rgb_at:
                assert np.array_equal(env.rgb(), ep["rgb"][rgb_at[t]])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "grouped_*.npz"))), ids=os.path.basename)
def test_grouped_trajectories(path):
    z = np.load(path)
    W, H, gravity, Q, use_features, _ = (int(v) for v in z["meta"])
    for ep in _episodes(z):
        env = OracleEnv(width=W, height=H, gravity=bool(gravity), queue_size=Q)
        env.set_sequence(ep["seq"])
        obs, _ = env.reset()
        T = len(ep["actions"])
        for t in range(T + 1):
            code = 0
            if t > 0:
                code, r, term, lines = env.grouped_step(int(ep["actions"][t - 1]), terminate_on_illegal=False)
                assert np.float32(r) == ep["reward"][t - 1] and term == bool(ep["terminated"][t - 1])
                assert lines == int(ep["lines"][t - 1])
                obs = env.obs()
            if use_features and code == 0:
                assert np.array_equal(env.features(obs), ep["info_board"][t])
            f, b, legal = env.grouped_observe(features=bool(use_features), boards=not use_features)
            got = f if use_features else b
            assert got.dtype == ep["obs"].dtype and np.array_equal(got, ep["obs"][t]), (path, t)
            assert np.array_equal(legal, ep["legal"][t])
            assert np.array_equal(env.board, ep["locked"][t])


def test_reference_known_answers():
    """The reference's own golden vectors: CSV placement, legal-mask table, mock-board
    features (tests/helpers/mock.py:35-47), reward 161 for a 4-line clear, seed-42 anchors."""
    k = np.load(os.path.join(GOLDEN, "reference_kat.npz"))
    # seed-42 anchors: reset(seed=42) through the numpy-exact 7-bag
    env = OracleEnv()
    obs, _ = env.reset(seed=42)
    assert np.array_equal(obs["board"], k["seed42_board"]) and np.array_equal(obs["queue"], k["seed42_queue"])
    assert np.array_equal(obs["holder"], k["seed42_holder"]) and np.array_equal(obs["active_tetromino_mask"], k["seed42_mask"])
    s = env.scalars()
    assert (s["active"], s["x"], s["y"], s["queue"]) == (3, 8, 0, [2, 6, 4, 1])
    # grouped fixture recipe (tests/test_grouped_env/conftest.py:16-33): mock board + vertical I
    env.board = k["mock_board"]
    env.set_active(0, rot=1)  # np.rot90(I) == one "clockwise" press
    _, boards, legal = env.grouped_observe(features=False, boards=True)
    assert np.array_equal(legal, k["legal_mask_vertical_i"])
    # the CSV was recorded with a RAW-id piece (cell value 1); ids 1 -> 2 inside the playfield
    exp = k["i_placement_csv"].copy()
    play = np.zeros_like(exp, bool); play[:20, 4:14] = True
    exp[play & (exp == 1)] = 2
    assert np.array_equal(boards[5 * 4 + 1], exp)
    for a in np.flatnonzero(legal == 0):
        assert np.all(boards[a] == 1)
    code, r, term, lines = env.grouped_step(5 * 4 + 1)
    assert code == 0 and np.array_equal(env.board, exp)
    # mock board features (tests/test_wrappers/test_feature_vector_observation.py)
    env2 = OracleEnv()
    env2.set_sequence(np.zeros(8, np.uint8)); env2.reset()
    f = env2.features({"board": k["mock_board"].copy(), "active_tetromino_mask": np.zeros_like(k["mock_board"])})
    assert np.array_equal(f[:10], k["mock_height"]) and f[10] == k["mock_max_height"][0]
    assert f[11] == k["mock_holes"][0] and f[12] == k["mock_bumpiness"][0]
    # tests/test_base_env/reward/test_base_env_line_clear.py:10-50: vertical I into a 4-row well -> 161
    env3 = OracleEnv(gravity=False)
    env3.set_sequence(np.zeros(8, np.uint8)); env3.reset()
    b = env3.board
    b[16:20, 4:13] = 2
    env3.board = b
    # vertical I (rot 1) occupies box column 1 -> board column x+1; the open well is padded column 13
    env3.set_active(0, rot=1, x=12, y=0)
    _, r, term, _, info = env3.step(5)
    assert (r, term, info["lines_cleared"]) == (161.0, False, 4)


def test_numpy_bag_stream():
    """PCG64 + masked-rejection Fisher-Yates restatement == numpy's own Generator.shuffle stream."""
    k = np.load(os.path.join(GOLDEN, "reference_kat.npz"))
    for seed in (42, 1, 7, 2**31 - 1, 123456789):
        rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        bag = np.arange(7, dtype=np.int8)
        rng.shuffle(bag)
        want, idx = [], 0
        for _ in range(700):
            want.append(int(bag[idx])); idx += 1
            if idx >= 7:
                rng.shuffle(bag); idx = 0
        env = OracleEnv()
        env.seed_numpy(seed)
        assert [int(v) for v in env.rnd_stream(700)] == want
        if seed == 42:
            assert want[:70] == [int(v) for v in k["seed42_stream"]]
    st = numpy_pcg64_state(42)
    assert st.dtype == np.uint64 and st.shape == (4,)


def test_numpy_true_randomizer_stream():
    """TrueRandomizer (components/tetromino_randomizer.py:105-136): `rng.integers(0, 7)` per draw.  The Lemire
    bounded-integer restatement (buffered next_uint32 halves of PCG64) == numpy's own Generator.integers stream,
    also when interleaved with nothing else (reset only reseeds)."""
    for seed in (42, 1, 7, 2**31 - 1, 123456789):
        rng = np.random.default_rng(seed)
        want = [int(rng.integers(0, 7)) for _ in range(2000)]
        env = OracleEnv()
        env.set_true_randomizer()
        env.seed_numpy(seed)
        assert [int(v) for v in env.rnd_stream(2000)] == want
    assert len(set(want)) == 7

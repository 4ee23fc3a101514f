"""Step-for-step check of the C oracle against the LIVE, unmodified reference NumPy env.

Build-container tool (needs /root/reference); run as `python -m oracle.validate_against_reference`.
It is also exercised by tests/test_oracle_vs_reference.py (skipped where the reference is absent).
Covers: base env (all 8 actions, gravity on/off, holder, queue sizes 4/5/7, default and wide
boards, seeded numpy 7-bag, seeded TrueRandomizer and injected streams, stepping after game over), the
FeatureVector / Rgb wrappers on the base env, and GroupedActionsObservations (boards, features,
legal mask, info["board"], illegal actions in both termination modes).
"""
import sys

import numpy as np

from . import _refload
from .tetris_oracle import OracleEnv


def _same_obs(a, b):
    return all(np.array_equal(a[k], b[k]) and a[k].dtype == b[k].dtype for k in ("board", "active_tetromino_mask", "holder", "queue"))


def check_base(ref, episodes, width, height, gravity, queue_size, seed0, injected, steps_after_over=3, max_steps=4000, true_random=False):
    R = ref
    rng = np.random.default_rng(seed0)
    n_steps = 0
    for ep in range(episodes):
        seq = rng.integers(0, 7, size=512) if injected else None
        env = R["make"](width=width, height=height, gravity=gravity, queue_size=queue_size, seq=seq, true_random=true_random)
        orc = OracleEnv(width=width, height=height, gravity=gravity, queue_size=queue_size)
        if true_random:
            orc.set_true_randomizer()
        rgbw = R["RgbObservation"](env)
        featw = R["FeatureVectorObservation"](env)
        if injected:
            orc.set_sequence(seq)
            o_ref, _ = env.reset()
            o_orc, _ = orc.reset()
        else:
            seed = int(rng.integers(1, 2**31))
            o_ref, _ = env.reset(seed=seed)
            o_orc, _ = orc.reset(seed=seed)
        assert _same_obs(o_ref, o_orc), "reset obs"
        over_left = steps_after_over
        for t in range(max_steps):
            a = int(rng.integers(0, 8))
            o_ref, r_ref, term_ref, trunc_ref, info_ref = env.step(a)
            o_orc, r_orc, term_orc, _, info_orc = orc.step(a)
            n_steps += 1
            assert _same_obs(o_ref, o_orc), f"obs ep{ep} t{t} a{a}"
            assert float(r_ref) == r_orc and bool(term_ref) == term_orc, f"reward/term ep{ep} t{t}"
            assert int(info_ref["lines_cleared"]) == info_orc["lines_cleared"]
            s = orc.scalars()
            assert (env.x, env.y, bool(env.has_swapped)) == (s["x"], s["y"], s["has_swapped"])
            assert np.array_equal(env.board, orc.board)
            assert np.array_equal(env.active_tetromino.matrix, orc.active_matrix())
            if t % 7 == 0:
                assert np.array_equal(rgbw.observation(o_ref), orc.rgb()), "rgb"
                f_ref = featw.observation({k: v.copy() for k, v in o_ref.items()})
                f_orc = orc.features({k: v.copy() for k, v in o_orc.items()})
                assert np.array_equal(f_ref, f_orc) and f_ref.dtype == f_orc.dtype, "features"
            if term_ref:
                over_left -= 1
                if over_left < 0:
                    break
    return n_steps


def check_holder(ref, episodes, holder_size, width, height, gravity, queue_size, seed0, max_steps=1500):
    """TetrominoHolder(size > 1) (components/tetromino_holder.py:14-57).  `Tetris(holder=...)` leaves self.holder unset in the
    reference (envs/tetris.py:140-145 only assigns the default), so the bigger holder is assigned after construction, like the
    scripted queue.  The reference's "holder" observation is RAGGED for a partly filled holder (np.hstack of the held pieces
    only); the oracle / CUDA env use the fixed shape (P, P * size): held pieces oldest first, ones in the empty slots.
    Checked: everything else bit for bit; holder[:, :P * held] == the reference's array and ones behind it; the RGB image
    (which pads the ragged array with ones itself, wrappers/observation.py:49-58) bit for bit."""
    R = ref
    rng = np.random.default_rng(seed0)
    n_steps = 0
    for ep in range(episodes):
        seq = rng.integers(0, 7, size=512)
        env = R["make"](width=width, height=height, gravity=gravity, queue_size=queue_size, seq=seq)
        env.holder = R["TetrominoHolder"](size=holder_size)
        orc = OracleEnv(width=width, height=height, gravity=gravity, queue_size=queue_size, holder_size=holder_size)
        orc.set_sequence(seq)
        rgbw = R["RgbObservation"](env)
        o_ref, _ = env.reset()
        o_orc, _ = orc.reset()
        for t in range(max_steps):
            a = int(rng.choice([0, 1, 2, 3, 4, 5, 5, 6, 6, 6, 7]))
            o_ref, r_ref, term_ref, _, info_ref = env.step(a)
            o_orc, r_orc, term_orc, _, info_orc = orc.step(a)
            n_steps += 1
            for k in ("board", "active_tetromino_mask", "queue"):
                assert np.array_equal(o_ref[k], o_orc[k]), f"{k} ep{ep} t{t} a{a}"
            held = len(env.holder.get_tetrominoes())
            assert held == orc.holder_len(), f"held ep{ep} t{t}"
            h_ref, h_orc = o_ref["holder"], o_orc["holder"]
            assert h_orc.shape == (4, 4 * holder_size)
            if held == 0:
                assert h_ref.shape == h_orc.shape and np.array_equal(h_ref, h_orc)
            else:
                assert h_ref.shape == (4, 4 * held) and np.array_equal(h_ref, h_orc[:, :4 * held]) and np.all(h_orc[:, 4 * held:] == 1), f"holder ep{ep} t{t}"
            assert float(r_ref) == r_orc and bool(term_ref) == term_orc and int(info_ref["lines_cleared"]) == info_orc["lines_cleared"]
            assert (env.x, env.y, bool(env.has_swapped)) == (orc.scalars()["x"], orc.scalars()["y"], orc.scalars()["has_swapped"])
            assert np.array_equal(env.active_tetromino.matrix, orc.active_matrix())
            if t % 5 == 0 and queue_size >= holder_size:
                assert np.array_equal(rgbw.observation(o_ref), orc.rgb()), f"rgb ep{ep} t{t}"
            if term_ref:
                break
    return n_steps


CUSTOM_SETS = {
    # a subset of the reference's pieces with other colours (bag of three)
    "iot": ([[[0, 0, 0, 0], [1, 1, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]], [[1, 1], [1, 1]], [[0, 1, 0], [1, 1, 1], [0, 0, 0]]],
            [[10, 200, 30], [250, 1, 99], [77, 77, 200]]),
    # five four-cell shapes that are not in the reference's set: a low I, an O in a corner of a 3 x 3 box, a T pointing down, a
    # skewed S, a long L (every column between a piece's first and last occupied one must hold a cell in each rotation, which is
    # true of every connected shape: the placement kernels describe a piece by its column profile)
    "odd5": ([[[0, 0, 0, 0], [0, 0, 0, 0], [1, 1, 1, 1], [0, 0, 0, 0]], [[1, 1, 0], [1, 1, 0], [0, 0, 0]], [[1, 1, 1], [0, 1, 0], [0, 0, 0]],
              [[0, 1, 0], [0, 1, 1], [0, 0, 1]], [[1, 0, 0, 0], [1, 0, 0, 0], [1, 1, 0, 0], [0, 0, 0, 0]]],
             [[1, 2, 3], [40, 50, 60], [200, 100, 0], [9, 99, 199], [255, 255, 255]]),
    # a single piece: the bag never shuffles
    "only_i": ([[[0, 0, 0, 0], [1, 1, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]]], [[0, 240, 240]]),
}


def check_custom_set(ref, name, episodes, width, height, gravity, queue_size, seed0, grouped, true_random=False, max_steps=1200):
    """Tetris(tetrominoes=[...]) (envs/tetris.py:88-89, 117-132): custom matrices / colours / set size, seeded numpy bag of
    len(set) pieces (or TrueRandomizer), base env + RGB image, or the grouped wrapper with features."""
    R = ref
    mats, cols = CUSTOM_SETS[name]
    rng = np.random.default_rng(seed0)
    n_steps = 0
    for ep in range(episodes):
        tets = [R["Tetromino"](i, list(c), np.array(m, dtype=np.uint8)) for i, (m, c) in enumerate(zip(mats, cols))]
        env = R["Tetris"](width=width, height=height, gravity=gravity, tetrominoes=tets)
        if true_random:
            env.randomizer = R["TrueRandomizer"](len(tets))
        env.queue = R["TetrominoQueue"](env.randomizer, size=queue_size)
        orc = OracleEnv(width=width, height=height, gravity=gravity, queue_size=queue_size)
        orc.set_tetrominoes(mats, cols)
        if true_random:
            orc.set_true_randomizer()
        seed = int(rng.integers(1, 2**31))
        if grouped:
            g = R["GroupedActionsObservations"](env, observation_wrappers=[R["FeatureVectorObservation"](env)])
            f_ref, info = g.reset(seed=seed)
            orc.reset(seed=seed)
            for t in range(max_steps):
                f_orc, _, legal = orc.grouped_observe()
                assert np.array_equal(np.asarray(f_ref, dtype=np.uint8), f_orc) and np.array_equal(np.asarray(info["action_mask"]).astype(np.uint8), legal), f"{name} grouped ep{ep} t{t}"
                a = int(rng.choice(np.flatnonzero(legal))) if rng.random() > 0.05 and legal.any() else int(rng.integers(0, 4 * width))
                f_ref, r_ref, term_ref, _, info = g.step(a)
                code, r_orc, term_orc, l_orc = orc.grouped_step(a)
                n_steps += 1
                assert float(r_ref) == r_orc and bool(term_ref) == term_orc and np.array_equal(env.board, orc.board), f"{name} grouped step ep{ep} t{t}"
                if term_ref:
                    break
            continue
        rgbw = R["RgbObservation"](env)
        o_ref, _ = env.reset(seed=seed)
        o_orc, _ = orc.reset(seed=seed)
        assert _same_obs(o_ref, o_orc), f"{name} reset"
        for t in range(max_steps):
            a = int(rng.integers(0, 8))
            o_ref, r_ref, term_ref, _, info_ref = env.step(a)
            o_orc, r_orc, term_orc, _, info_orc = orc.step(a)
            n_steps += 1
            assert _same_obs(o_ref, o_orc), f"{name} obs ep{ep} t{t} a{a}"
            assert float(r_ref) == r_orc and bool(term_ref) == term_orc and int(info_ref["lines_cleared"]) == info_orc["lines_cleared"]
            assert np.array_equal(env.board, orc.board) and np.array_equal(env.active_tetromino.matrix, orc.active_matrix())
            if t % 6 == 0:
                assert np.array_equal(rgbw.observation(o_ref), orc.rgb()), f"{name} rgb"
            if term_ref:
                break
    return n_steps


def check_grouped(ref, episodes, width, height, gravity, queue_size, seed0, use_features, terminate_on_illegal, max_steps=600, greedy=False):
    from .make_golden import greedy_action

    R = ref
    rng = np.random.default_rng(seed0)
    n_steps = 0
    for ep in range(episodes):
        seq = rng.integers(0, 7, size=512)
        base = R["make"](width=width, height=height, gravity=gravity, queue_size=queue_size, seq=seq)
        wrappers = [R["FeatureVectorObservation"](base)] if use_features else None
        env = R["GroupedActionsObservations"](base, observation_wrappers=wrappers, terminate_on_illegal_action=terminate_on_illegal)
        orc = OracleEnv(width=width, height=height, gravity=gravity, queue_size=queue_size)
        orc.set_sequence(seq)
        g_ref, info = env.reset()
        o0, _ = orc.reset()
        ib = orc.features(o0) if use_features else None
        f, b, legal = orc.grouped_observe(features=use_features, boards=not use_features)
        g_orc = f if use_features else b
        assert np.array_equal(g_ref, g_orc) and g_ref.dtype == g_orc.dtype, "grouped reset obs"
        assert np.array_equal(info["action_mask"].astype(np.uint8), legal)
        if use_features:
            assert np.array_equal(info["board"], ib)
        for t in range(max_steps):
            # mostly legal actions, sometimes any action (to hit the illegal path)
            if greedy and use_features and rng.random() > 0.05:
                a = greedy_action(g_ref, legal, width)
            elif rng.random() < 0.1:
                a = int(rng.integers(0, 4 * width))
            else:
                a = int(rng.choice(np.flatnonzero(legal)))
            g_ref, r_ref, term_ref, _, info = env.step(a)
            code, r_orc, term_orc, l_orc = orc.grouped_step(a, terminate_on_illegal)
            n_steps += 1
            assert float(r_ref) == r_orc and bool(term_ref) == term_orc, f"grouped reward/term ep{ep} t{t} a{a}"
            assert int(info["lines_cleared"]) == l_orc
            if code == 1:
                # illegal + terminate: reference returns ones * high (float), env untouched
                assert np.all(g_ref == float(height * width))
                break
            o = orc.obs()
            if code == 0 and use_features:
                assert np.array_equal(info["board"], orc.features(o)), "info board"
            elif code == 0:
                assert _same_obs(info["board"], o), "info board dict"
            f, b, legal = orc.grouped_observe(features=use_features, boards=not use_features)
            g_orc = f if use_features else b
            assert np.array_equal(g_ref, g_orc), f"grouped obs ep{ep} t{t}"
            assert np.array_equal(info["action_mask"].astype(np.uint8), legal)
            assert np.array_equal(base.board, orc.board)
            if term_ref:
                break
    return n_steps


def check_grouped_constructed(ref, n_boards, width, height, queue_size, seed0, use_features):
    """GroupedActionsObservations on CONSTRUCTED boards poked into the live reference env the way its own tests do
    (env.unwrapped.board / active_tetromino): nearly full bottom rows with wells (placements that clear 1-4 rows), rubble,
    and towers into the spawn rows (placements ending in row 0, game-over placements) -- the rare situations of the placement
    enumeration.  Every placement's board or feature row, the legal mask and one executed placement."""
    import copy

    R = ref
    rng = np.random.default_rng(seed0)
    W, H = width, height
    n_checked = cleared = 0
    for i in range(n_boards):
        seq = rng.integers(0, 7, size=64)
        base = R["make"](width=W, height=H, gravity=False, queue_size=queue_size, seq=seq)
        wrappers = [R["FeatureVectorObservation"](base)] if use_features else None
        env = R["GroupedActionsObservations"](base, observation_wrappers=wrappers)
        orc = OracleEnv(width=W, height=H, gravity=False, queue_size=queue_size)
        orc.set_sequence(seq)
        env.reset()
        orc.reset()
        b = orc.board
        k = int(rng.integers(1, 6))
        b[H - k:H, 4:4 + W] = rng.integers(2, 9, size=(k, W))
        wells = rng.choice(W, size=int(rng.integers(1, 3)), replace=False)
        depth = int(rng.integers(1, k + 1))
        b[H - k:H - k + depth, 4 + wells] = 0
        noise = rng.random((3, W)) < 0.3
        b[H - k - 3:H - k, 4:4 + W] = np.where(noise, rng.integers(2, 9, size=(3, W)), 0)
        if i % 3 == 0:
            for c in rng.choice(W, size=int(rng.integers(1, W)), replace=False):
                top = int(rng.integers(0, 6))
                col = np.where(rng.random(H - k - top) < 0.8, rng.integers(2, 9, size=H - k - top), 0)
                col[0] = 3
                b[top:H - k, 4 + c] = col
        piece, rot = int(rng.integers(0, 7)), int(rng.integers(0, 4))
        orc.board = b
        orc.set_active(piece, rot)
        base.board = b.copy()
        t = copy.deepcopy(base.tetrominoes[piece])
        for _ in range(rot):
            t = base.rotate(t, True)
        base.active_tetromino = t
        g_ref = env.observation(base._get_obs())
        f, bb, legal = orc.grouped_observe(features=use_features, boards=not use_features)
        g_orc = f if use_features else bb
        assert np.array_equal(g_ref, g_orc), f"constructed board {i}: grouped obs"
        assert np.array_equal(env.legal_actions_mask.astype(np.uint8), legal), f"constructed board {i}: legal mask"
        _, _, ln = orc.grouped_observe_lines()
        cleared += int((ln > 0).sum())
        a = int(rng.choice(np.flatnonzero(legal)))
        g_ref, r_ref, term_ref, _, info = env.step(a)
        code, r_orc, term_orc, l_orc = orc.grouped_step(a, True)
        assert float(r_ref) == r_orc and bool(term_ref) == term_orc and int(info["lines_cleared"]) == l_orc, f"constructed board {i}: step"
        assert np.array_equal(base.board, orc.board)
        f, bb, legal = orc.grouped_observe(features=use_features, boards=not use_features)
        assert np.array_equal(g_ref, f if use_features else bb), f"constructed board {i}: obs after the step"
        n_checked += 1
    assert cleared > n_boards, "the constructed boards should produce many row-clearing placements"
    return n_checked


def run(scale=1):
    ref = _refload.load()
    n = 0
    n += check_base(ref, 6 * scale, 10, 20, True, 4, 1, injected=False)
    n += check_base(ref, 6 * scale, 10, 20, True, 4, 2, injected=True)
    n += check_base(ref, 4 * scale, 10, 20, False, 7, 3, injected=True, max_steps=1500)
    n += check_base(ref, 3 * scale, 20, 40, True, 5, 4, injected=True)
    n += check_base(ref, 3 * scale, 6, 8, True, 4, 5, injected=False)
    n += check_base(ref, 3 * scale, 13, 9, True, 3, 6, injected=True)
    n += check_base(ref, 3 * scale, 10, 20, True, 4, 7, injected=False, true_random=True)
    n += check_base(ref, 2 * scale, 10, 20, False, 7, 8, injected=False, true_random=True, max_steps=1200)
    n += check_holder(ref, 3 * scale, 2, 10, 20, True, 4, 31)
    n += check_holder(ref, 2 * scale, 3, 10, 20, False, 4, 32, max_steps=800)
    n += check_holder(ref, 2 * scale, 4, 10, 20, True, 7, 33)
    n += check_holder(ref, 1 * scale, 2, 20, 40, True, 5, 34)
    n += check_custom_set(ref, "iot", 3 * scale, 10, 20, True, 4, 41, grouped=False)
    n += check_custom_set(ref, "odd5", 3 * scale, 10, 20, True, 4, 42, grouped=False)
    n += check_custom_set(ref, "odd5", 2 * scale, 12, 16, False, 7, 43, grouped=False, true_random=True)
    n += check_custom_set(ref, "only_i", 1 * scale, 10, 20, True, 4, 44, grouped=False)
    g = 0
    g += check_custom_set(ref, "odd5", 2 * scale, 10, 20, False, 4, 45, grouped=True, max_steps=300)
    g += check_custom_set(ref, "iot", 2 * scale, 10, 20, False, 4, 46, grouped=True, max_steps=300)
    g += check_grouped(ref, 4 * scale, 10, 20, False, 4, 11, True, True)
    g += check_grouped(ref, 3 * scale, 10, 20, False, 4, 12, False, False)
    g += check_grouped(ref, 3 * scale, 10, 20, True, 4, 13, True, False)
    g += check_grouped(ref, 2 * scale, 20, 40, False, 5, 14, True, True, max_steps=300)
    g += check_grouped(ref, 2 * scale, 7, 10, False, 4, 15, False, True)
    g += check_grouped_constructed(ref, 60 * scale, 10, 20, 4, 21, True)
    g += check_grouped_constructed(ref, 30 * scale, 10, 20, 4, 22, False)
    g += check_grouped_constructed(ref, 16 * scale, 20, 40, 5, 23, True)
    g += check_grouped(ref, 2 * scale, 10, 20, False, 4, 16, True, True, max_steps=400, greedy=True)
    g += check_grouped(ref, 1 * scale, 20, 40, False, 5, 17, True, False, max_steps=250, greedy=True)
    return n, g


if __name__ == "__main__":
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n, g = run(scale)
    print(f"oracle == reference on {n} base steps and {g} grouped steps")

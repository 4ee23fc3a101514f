"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Every array is produced by importing /root/reference through oracle/_refload.py (gymnasium
stand-in) and driving it with recorded piece streams / seeds and action streams.  The fixtures
travel to the GPU box; the reference does not.  Files:

  base_<cfg>.npz      per-step obs dict, reward, terminated, lines, x, y for random-action episodes
  grouped_<cfg>.npz   per-step grouped obs (features or boards), legal mask, info["board"], reward, ...
  rgb_<cfg>.npz       RgbObservation frames sampled along a base episode
  reference_kat.npz   the reference's own known-answer vectors (tests/test_grouped_env/
                      expected_result_i_placement.csv, the legal-mask table, the mock board + features)
"""
import os

import numpy as np

from . import _refload

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def record_base(R, name, width, height, gravity, queue_size, seeds, injected, max_steps, after_over=2, rgb_every=0):
    eps = []
    for s in seeds:
        rng = np.random.default_rng(1000 + s)
        # injected: 0 = seeded numpy 7-bag, 1 = injected piece stream, 2 = seeded TrueRandomizer
        seq = rng.integers(0, 7, size=256).astype(np.uint8) if injected == 1 else np.zeros(0, np.uint8)
        env = R["make"](width=width, height=height, gravity=gravity, queue_size=queue_size, seq=seq if injected == 1 else None,
                        true_random=injected == 2)
        rgbw = R["RgbObservation"](env)
        featw = R["FeatureVectorObservation"](env)
        obs, _ = env.reset(seed=None if injected == 1 else s)
        rec = dict(board=[obs["board"]], mask=[obs["active_tetromino_mask"]], holder=[obs["holder"]], queue=[obs["queue"]],
                   reward=[], terminated=[], lines=[], actions=[], x=[env.x], y=[env.y], rgb=[], rgb_t=[], feat=[], locked=[env.board.copy()])
        rec["feat"].append(featw.observation({k: v.copy() for k, v in obs.items()}))
        left = after_over
        for t in range(max_steps):
            a = int(rng.integers(0, 8))
            obs, r, term, _, info = env.step(a)
            rec["actions"].append(a)
            rec["board"].append(obs["board"]); rec["mask"].append(obs["active_tetromino_mask"])
            rec["holder"].append(obs["holder"]); rec["queue"].append(obs["queue"])
            rec["reward"].append(np.float32(r)); rec["terminated"].append(bool(term)); rec["lines"].append(int(info["lines_cleared"]))
            rec["x"].append(env.x); rec["y"].append(env.y); rec["locked"].append(env.board.copy())
            rec["feat"].append(featw.observation({k: v.copy() for k, v in obs.items()}))
            if rgb_every and t % rgb_every == 0:
                rec["rgb"].append(rgbw.observation(obs)); rec["rgb_t"].append(t + 1)
            if term:
                left -= 1
                if left < 0:
                    break
        eps.append((s, seq, rec))
    out = {"meta": np.array([width, height, int(gravity), queue_size, int(injected), len(eps)], np.int64)}
    for i, (s, seq, rec) in enumerate(eps):
        out[f"e{i}_seed"] = np.int64(s)
        out[f"e{i}_seq"] = seq
        for k, v in rec.items():
            if k in ("rgb", "rgb_t") and not v:
                continue
            out[f"e{i}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, f"base_{name}.npz"), **out)
    return sum(len(r["actions"]) for _, _, r in eps)


def greedy_action(feats, legal, width):
    """Deterministic line-clearing policy used to make goldens with many cleared rows: among legal
    placements minimise (holes, bumpiness, max height), lowest index wins ties."""
    f = feats.astype(np.int64)
    cost = f[:, width + 1] * 10000 + f[:, width + 2] * 100 + f[:, width]
    cost = np.where(legal > 0, cost, np.iinfo(np.int64).max)
    return int(np.argmin(cost))


def record_grouped(R, name, width, height, gravity, queue_size, seeds, use_features, max_steps, greedy=False):
    eps = []
    for s in seeds:
        rng = np.random.default_rng(2000 + s)
        seq = rng.integers(0, 7, size=256).astype(np.uint8)
        base = R["make"](width=width, height=height, gravity=gravity, queue_size=queue_size, seq=seq)
        wr = [R["FeatureVectorObservation"](base)] if use_features else None
        env = R["GroupedActionsObservations"](base, observation_wrappers=wr, terminate_on_illegal_action=False)
        g, info = env.reset()
        rec = dict(obs=[g], legal=[info["action_mask"].astype(np.uint8)], reward=[], terminated=[], lines=[], actions=[], locked=[base.board.copy()])
        if use_features:
            rec["info_board"] = [info["board"]]
        for t in range(max_steps):
            legal = np.flatnonzero(info["action_mask"])
            if greedy and rng.random() > 0.05:
                a = greedy_action(g, info["action_mask"], width)
            else:
                a = int(rng.integers(0, 4 * width)) if rng.random() < 0.08 else int(rng.choice(legal))
            g, r, term, _, info = env.step(a)
            rec["actions"].append(a); rec["obs"].append(g); rec["legal"].append(info["action_mask"].astype(np.uint8))
            rec["reward"].append(np.float32(r)); rec["terminated"].append(bool(term)); rec["lines"].append(int(info["lines_cleared"]))
            rec["locked"].append(base.board.copy())
            if use_features:
                # illegal (no-op) steps carry no info["board"]: repeat the previous one as a placeholder
                rec["info_board"].append(info["board"] if "board" in info else rec["info_board"][-1])
            if term:
                break
        eps.append((s, seq, rec))
    out = {"meta": np.array([width, height, int(gravity), queue_size, int(use_features), len(eps)], np.int64)}
    for i, (s, seq, rec) in enumerate(eps):
        out[f"e{i}_seq"] = seq
        for k, v in rec.items():
            out[f"e{i}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, f"grouped_{name}.npz"), **out)
    return sum(len(r["actions"]) for _, _, r in eps)


def record_o_script(R):
    """Base env, gravity on, all-O stream; actions walk each O to a column then hard-drop so that
    rows fill and clear (exercises clear_filled_rows + reward lines^2*W + alife through Tetris.step)."""
    seq = np.ones(64, np.uint8)
    env = R["make"](seq=seq)
    obs, _ = env.reset()
    rec = dict(board=[obs["board"]], mask=[obs["active_tetromino_mask"]], holder=[obs["holder"]], queue=[obs["queue"]],
               reward=[], terminated=[], lines=[], actions=[], x=[env.x], y=[env.y], locked=[env.board.copy()])
    cols = [4, 6, 8, 10, 12] * 6  # padded x targets; O spawns at x=8
    for tx in cols:
        acts = []
        dx = tx - 8
        acts += [0 if dx < 0 else 1] * abs(dx)
        acts += [5]
        for a in acts:
            obs, r, term, _, info = env.step(a)
            rec["actions"].append(a)
            rec["board"].append(obs["board"]); rec["mask"].append(obs["active_tetromino_mask"])
            rec["holder"].append(obs["holder"]); rec["queue"].append(obs["queue"])
            rec["reward"].append(np.float32(r)); rec["terminated"].append(bool(term)); rec["lines"].append(int(info["lines_cleared"]))
            rec["x"].append(env.x); rec["y"].append(env.y); rec["locked"].append(env.board.copy())
    out = {"meta": np.array([10, 20, 1, 4, 1, 1], np.int64), "e0_seed": np.int64(0), "e0_seq": seq}
    for k, v in rec.items():
        out[f"e0_{k}"] = np.asarray(v)
    assert sum(rec["lines"]) >= 10, sum(rec["lines"])
    np.savez_compressed(os.path.join(OUT, "base_d_o_script.npz"), **out)
    return len(rec["actions"])


def record_kat(R):
    """The reference's own golden vectors, re-derived by running its fixtures' recipe."""
    ref_tests = os.path.join(_refload.REF, "tests")
    csv = np.genfromtxt(os.path.join(ref_tests, "test_grouped_env", "expected_result_i_placement.csv"), delimiter=",").astype(np.uint8)
    # tests/helpers/mock.py:5-47 (generate_example_board_with_features), executed from the reference tree
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_mock", os.path.join(ref_tests, "helpers", "mock.py"))
    mock = importlib.util.module_from_spec(spec); spec.loader.exec_module(mock)
    env = R["make"]()
    env.reset(seed=42)
    board, height, max_height, holes, bump = mock.generate_example_board_with_features(env)
    # tests/test_grouped_env/actions/test_grouped_actions.py:18-26
    legal = np.array([[0, 1, 1, 1, 1, 1, 1, 1, 1, 1], [0, 0, 1, 1, 1, 1, 1, 1, 1, 0],
                      [1, 1, 1, 1, 1, 1, 1, 1, 1, 1], [0, 0, 1, 1, 1, 1, 1, 1, 1, 0]], np.uint8).reshape(40, order="F")
    # seed-42 anchors, taken from the live reference
    obs, _ = env.reset(seed=42)
    stream = []
    rnd = R["Tetris"]().randomizer
    rnd.reset(seed=42)
    for _ in range(70):
        stream.append(int(rnd.get_next_tetromino()))
    np.savez_compressed(os.path.join(OUT, "reference_kat.npz"), i_placement_csv=csv, mock_board=board,
                        mock_height=height, mock_max_height=max_height, mock_holes=holes, mock_bumpiness=bump,
                        legal_mask_vertical_i=legal, seed42_board=obs["board"], seed42_queue=obs["queue"],
                        seed42_holder=obs["holder"], seed42_mask=obs["active_tetromino_mask"],
                        seed42_stream=np.array(stream, np.uint8))


def main():
    os.makedirs(OUT, exist_ok=True)
    R = _refload.load()
    n = 0
    n += record_base(R, "d_seeded", 10, 20, True, 4, [42, 7, 123], injected=False, max_steps=1200, rgb_every=25)
    n += record_base(R, "d_injected_q7", 10, 20, True, 7, [1, 2], injected=True, max_steps=1200)
    n += record_base(R, "d_nogravity", 10, 20, False, 4, [3], injected=True, max_steps=600)
    n += record_base(R, "x_wide_q5", 20, 40, True, 5, [5, 6], injected=True, max_steps=1500, rgb_every=40)
    n += record_base(R, "odd_13x9_q3", 13, 9, True, 3, [8], injected=True, max_steps=400)
    n += record_o_script(R)
    n += record_base(R, "d_true_randomizer", 10, 20, True, 4, [42, 9], injected=2, max_steps=900)
    g = 0
    g += record_grouped(R, "d_features_greedy", 10, 20, False, 4, [9, 10], True, 400, greedy=True)
    g += record_grouped(R, "x_features_greedy_q5", 20, 40, False, 5, [11], True, 250, greedy=True)
    g += record_grouped(R, "d_features", 10, 20, False, 4, [1, 2, 3], True, 200)
    g += record_grouped(R, "d_boards", 10, 20, False, 4, [4], False, 200)
    g += record_grouped(R, "d_features_gravity", 10, 20, True, 4, [5], True, 200)
    g += record_grouped(R, "x_features_q5", 20, 40, False, 5, [6], True, 120)
    record_kat(R)
    print(f"recorded {n} base steps, {g} grouped steps into {OUT}")


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "true_randomizer":   # add just this fixture, leave the others untouched
        print(record_base(_refload.load(), "d_true_randomizer", 10, 20, True, 4, [42, 9], injected=2, max_steps=900), "steps")
    else:
        main()

"""GPU parity of the fused K-step heuristic rollout (tg_rollout) against the same integer policy driven
step by step through the C oracle's grouped enumeration."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

WEIGHTS = (-51, 76, -36, -18)


def oracle_policy_step(o, W, weights):
    feats, legal, lines = o.grouped_observe_lines()
    f = feats.astype(np.int64)
    score = weights[0] * f[:, :W].sum(1) + weights[1] * lines + weights[2] * f[:, W + 1] + weights[3] * f[:, W + 2]
    ok = lines >= 0
    if ok.any():
        s = np.where(ok, score, np.iinfo(np.int64).min)
        a = int(np.argmax(s))          # first maximum = lowest index
    else:
        a = int(np.flatnonzero(legal)[0])
    return a


@pytest.mark.parametrize("cfg", [
    dict(width=10, height=20, queue_size=7, autoreset="next_step", K=120),
    dict(width=10, height=20, queue_size=4, autoreset="disabled", K=60),
    dict(width=20, height=40, queue_size=5, autoreset="same_step", K=150),
    dict(width=6, height=12, queue_size=3, autoreset="next_step", K=200),
], ids=lambda c: f"{c['width']}x{c['height']}-{c['autoreset']}")
def test_rollout_matches_stepwise_oracle_policy(cfg):
    from gpu_util import np_
    from oracle.tetris_oracle import OracleEnv
    from tetris_gymnasium_b200.envs.tetris import Tetris

    W, H, Q, K, mode = cfg["width"], cfg["height"], cfg["queue_size"], cfg["K"], cfg["autoreset"]
    n, L = 150, 83
    rng = np.random.default_rng(W + H)
    seqs = rng.integers(0, 7, size=(n, L)).astype(np.uint8)
    env = Tetris(width=W, height=H, gravity=False, queue_size=Q, num_envs=n, randomizer_mode="sequence",
                 piece_sequences=seqs, autoreset_mode=mode)
    env.reset()
    orcs = [OracleEnv(width=W, height=H, gravity=False, queue_size=Q) for _ in range(n)]
    ep = lines_tot = 0
    ret = 0.0
    length = 0
    for i, o in enumerate(orcs):
        o.set_sequence(seqs[i])
        o.reset()
    # two launches (K1 + K2) must equal one trajectory of K steps: state carries over between launches
    K1 = K // 3
    env.rollout(WEIGHTS, K1)
    last = env.rollout(WEIGHTS, K - K1, trace=True)
    want_last = np.full(n, -1)
    for i, o in enumerate(orcs):
        pending = False
        er, el, eln = 0.0, 0, 0
        for t in range(K):
            if mode == "next_step" and pending:
                o.reset(); pending = False; want_last[i] = -1
                continue
            a = oracle_policy_step(o, W, WEIGHTS)
            want_last[i] = a
            code, r, term, ln = o.grouped_step(a)
            assert code == 0
            er += r; el += 1; eln += ln
            if term:
                ep += 1; ret += er; length += el; lines_tot += eln
                er, el, eln = 0.0, 0, 0
                if mode == "next_step":
                    pending = True
                elif mode == "same_step":
                    o.reset()
    st = env.get_state()
    assert np.array_equal(np_(st["board"]), np.stack([o.board for o in orcs]))
    sc = [o.scalars() for o in orcs]
    assert np.array_equal(np_(st["piece"]), np.array([s["active"] for s in sc]))
    assert np.array_equal(np_(st["queue"]), np.array([s["queue"] for s in sc]))
    assert np.array_equal(np_(last), want_last)
    stats = {k: float(v) for k, v in env.episode_stats().items()}
    assert stats["episodes"] == ep and stats["sum_length"] == length and stats["sum_lines"] == lines_tot
    assert abs(stats["sum_return"] - ret) < 1e-6 * max(1.0, abs(ret))

"""GPU parity of the grouped path (tg_grouped_observe / tg_grouped_step) and the feature / RGB wrappers
against the golden fixtures from the unmodified reference and against the C oracle."""
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


def _make(W, H, gravity, Q, n, seqs, use_features, terminate, autoreset="disabled"):
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations

    base = Tetris(width=W, height=H, gravity=bool(gravity), queue_size=Q, num_envs=n, randomizer_mode="sequence",
                  piece_sequences=seqs, autoreset_mode=autoreset, terminate_on_illegal_action=terminate)
    wr = [FeatureVectorObservation(base)] if use_features else None
    return base, GroupedActionsObservations(base, observation_wrappers=wr, terminate_on_illegal_action=terminate)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "grouped_*.npz"))), ids=os.path.basename)
def test_golden_grouped_trajectories(path):
    from gpu_util import np_

    z = np.load(path)
    W, H, gravity, Q, use_features, _ = (int(v) for v in z["meta"])
    for ep in _episodes(z):
        base, env = _make(W, H, gravity, Q, 1, ep["seq"][None, :], bool(use_features), False)
        g, info = env.reset()
        T = len(ep["actions"])
        for t in range(T + 1):
            legal_action = True
            if t > 0:
                a = int(ep["actions"][t - 1])
                legal_action = bool(ep["legal"][t - 1][a])
                g, r, term, trunc, info = env.step(torch.tensor([a]))
                assert np_(r)[0] == ep["reward"][t - 1], (path, t)
                assert bool(np_(term)[0]) == bool(ep["terminated"][t - 1])
                assert int(np_(info["lines_cleared"])[0]) == int(ep["lines"][t - 1])
            got = np_(g)[0]
            assert got.dtype == ep["obs"].dtype and np.array_equal(got, ep["obs"][t]), (path, t)
            assert np.array_equal(np_(info["action_mask"])[0], ep["legal"][t]), (path, t)
            if use_features and legal_action:
                assert np.array_equal(np_(info["board"])[0], ep["info_board"][t]), (path, t)
            assert np.array_equal(np_(base.get_state()["board"])[0], ep["locked"][t]), (path, t)
        base.close()


def test_reference_known_answers_on_gpu():
    """The reference's own golden vectors through the CUDA path: CSV placement (action 21 on the mock
    board with a vertical I), legal-mask table, illegal -> ones, mock-board features, reward 161."""
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations
    from gpu_util import np_

    k = np.load(os.path.join(GOLDEN, "reference_kat.npz"))
    base = Tetris(num_envs=1, randomizer_mode="numpy", autoreset_mode="disabled")
    env = GroupedActionsObservations(base)
    env.reset(seed=42)
    base.set_state(board=k["mock_board"], piece=0, rotation=1)   # vertical I = np.rot90(I)
    boards = np_(env.observation())[0]
    legal = np_(env.legal_actions_mask)[0]
    assert np.array_equal(legal, k["legal_mask_vertical_i"])
    exp = k["i_placement_csv"].copy()          # recorded with a raw-id piece (value 1): 1 -> 2 in the playfield
    play = np.zeros_like(exp, bool); play[:20, 4:14] = True
    exp[play & (exp == 1)] = 2
    assert np.array_equal(boards[5 * 4 + 1], exp)
    for a in np.flatnonzero(legal == 0):
        assert (boards[a] == 1).all()
    env.step(torch.tensor([5 * 4 + 1]))
    assert np.array_equal(np_(base.get_state()["board"])[0], exp)
    # mock board features (tests/helpers/mock.py:35-47) through the grouped feature path: zeros board trick --
    # put the mock board in, make every placement game-over-free and read info-style features via tg_features
    base2 = Tetris(num_envs=1, randomizer_mode="numpy", autoreset_mode="disabled")
    base2.reset(seed=42)
    fw = FeatureVectorObservation(base2)
    base2.set_state(board=k["mock_board"], y=0)
    f = np_(fw.observation())[0]
    assert np.array_equal(f[:10], k["mock_height"]) and f[10] == k["mock_max_height"][0]
    assert f[11] == k["mock_holes"][0] and f[12] == k["mock_bumpiness"][0]
    # 4-line clear with a vertical I -> reward 161 (tests/test_base_env/reward/test_base_env_line_clear.py:10-50)
    base3 = Tetris(num_envs=1, gravity=False, randomizer_mode="numpy", autoreset_mode="disabled")
    base3.reset(seed=42)
    b = np_(base3.get_state()["board"])[0].copy()
    b[16:20, 4:13] = 2
    base3.set_state(board=b, piece=0, rotation=1, x=12, y=0)
    _, r, term, _, info = base3.step(torch.tensor([5]))
    assert (float(r[0]), bool(term[0]), int(info["lines_cleared"][0])) == (161.0, False, 4)
    # game-over placements are all zeros (tests/test_grouped_env/observation/test_grouped_observations.py:44-68)
    full = k["mock_board"].copy(); full[0:20, 4:14] = 2
    base.set_state(board=full, piece=0, rotation=1, x=7, y=0)
    boards = np_(env.observation())[0]
    legal = np_(env.legal_actions_mask)[0]
    assert any(legal[i] == 1 and (boards[i] == 0).all() for i in range(40))


@pytest.mark.parametrize("cfg", [
    dict(W=10, H=20, gravity=False, Q=4, feats=True, terminate=True),
    dict(W=10, H=20, gravity=False, Q=4, feats=False, terminate=False),
    dict(W=10, H=20, gravity=True, Q=7, feats=True, terminate=False),
    dict(W=20, H=40, gravity=False, Q=5, feats=True, terminate=True),
    dict(W=20, H=40, gravity=False, Q=5, feats=False, terminate=True),
    dict(W=7, H=10, gravity=False, Q=2, feats=True, terminate=True),
    dict(W=13, H=30, gravity=True, Q=3, feats=False, terminate=False),
], ids=lambda c: f"{c['W']}x{c['H']}g{int(c['gravity'])}f{int(c['feats'])}t{int(c['terminate'])}")
def test_grouped_random_and_greedy_vs_oracle(cfg):
    """Batched grouped env (NEXT_STEP autoreset) vs n oracle envs: half the envs play a line-clearing greedy
    policy, the rest random placements with occasional illegal actions."""
    from gpu_util import np_
    from oracle.tetris_oracle import OracleEnv

    W, H, Q = cfg["W"], cfg["H"], cfg["Q"]
    n, T, L = 77, 150 if W <= 13 else 90, 61
    A = 4 * W
    rng = np.random.default_rng(W * 100 + H)
    seqs = rng.integers(0, 7, size=(n, L)).astype(np.uint8)
    base, env = _make(W, H, cfg["gravity"], Q, n, seqs, cfg["feats"], cfg["terminate"], autoreset="next_step")
    orcs = [OracleEnv(width=W, height=H, gravity=cfg["gravity"], queue_size=Q) for _ in range(n)]
    for i, o in enumerate(orcs):
        o.set_sequence(seqs[i])

    def orc_observe():
        res = [o.grouped_observe(features=cfg["feats"], boards=not cfg["feats"]) for o in orcs]
        g = np.stack([r[0] if cfg["feats"] else r[1] for r in res])
        return g, np.stack([r[2] for r in res])

    g, info = env.reset()
    ib = []
    for o in orcs:
        ob, _ = o.reset()
        ib.append(o.features(ob))
    g2, legal2 = orc_observe()
    assert np.array_equal(np_(g), g2) and np.array_equal(np_(info["action_mask"]), legal2)
    if cfg["feats"]:
        assert np.array_equal(np_(info["board"]), np.stack(ib))
    pending = np.zeros(n, bool)
    total_lines = 0
    for t in range(T):
        legal = legal2
        a = np.empty(n, np.int64)
        for i in range(n):
            if cfg["feats"] and i % 2 == 0 and rng.random() > 0.05:
                f = g2[i].astype(np.int64)
                cost = f[:, W + 1] * 10000 + f[:, W + 2] * 100 + f[:, W]
                cost = np.where(legal[i] > 0, cost, np.iinfo(np.int64).max)
                a[i] = int(np.argmin(cost))
            elif rng.random() < 0.06:
                a[i] = rng.integers(0, A)
            else:
                a[i] = rng.choice(np.flatnonzero(legal[i]))
        g, r, term, trunc, info = env.step(torch.from_numpy(a))
        r2, t2, l2 = np.zeros(n, np.float32), np.zeros(n, bool), np.zeros(n, np.int32)
        high_rows = np.zeros(n, bool)
        ibs = [None] * n
        for i, o in enumerate(orcs):
            if pending[i]:
                ob, _ = o.reset()
                ibs[i] = o.features(ob) if cfg["feats"] else None
                continue
            code, rr, tt, ll = o.grouped_step(int(a[i]), cfg["terminate"])
            r2[i], t2[i], l2[i] = rr, tt, ll
            high_rows[i] = code == 1
            if cfg["feats"] and code == 0:
                ibs[i] = o.features(o.obs())
        pending = t2.copy()
        g2n, legal2n = orc_observe()
        # illegal + terminate: obs is filled with `high`, mask unchanged, env untouched
        for i in np.flatnonzero(high_rows):
            g2n[i] = np.uint8(min(H * W, 255))      # observation_space.high = H * W, saturated to the uint8 range
            legal2n[i] = legal[i]
        assert np.array_equal(np_(r), r2), t
        assert np.array_equal(np_(term), t2) and np.array_equal(np_(info["lines_cleared"]), l2)
        assert np.array_equal(np_(info["action_mask"]), legal2n), t
        if not np.array_equal(np_(g), g2n):
            bad = np.flatnonzero((np_(g) != g2n).reshape(n, -1).any(1))
            raise AssertionError(f"t={t} grouped obs differs for envs {bad[:5]}")
        if cfg["feats"]:
            ibg = np_(info["board"])
            for i in range(n):
                if ibs[i] is not None:
                    assert np.array_equal(ibg[i], ibs[i]), (t, i)
        g2, legal2 = g2n, legal2n
        total_lines += int(l2.sum())
    if cfg["feats"]:
        assert total_lines > 20
    assert np.array_equal(np_(base.get_state()["board"]), np.stack([o.board for o in orcs]))


@pytest.mark.parametrize("W,H", [(10, 20), (20, 40), (13, 30), (24, 28)], ids=lambda v: str(v))
def test_grouped_boards_with_line_clears_vs_oracle(W, H):
    """Board-image enumeration (no observation wrappers) on constructed boards whose bottom rows are full but for
    one or two wells, so that many of the 4W placements clear 1-4 rows: every placement image, the legal mask and the
    executed placement vs the oracle.  (10,20)/(20,40)/(24,28) take the streaming kernel (image bytes % 16 == 0),
    (13,30) the generic one."""
    from gpu_util import np_
    from oracle.tetris_oracle import OracleEnv
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import GroupedActionsObservations

    n, A = 96, 4 * W
    rng = np.random.default_rng(7 * W + H)
    seqs = rng.integers(0, 7, size=(n, 31)).astype(np.uint8)
    base = Tetris(width=W, height=H, gravity=False, queue_size=3, num_envs=n, randomizer_mode="sequence",
                  piece_sequences=seqs, autoreset_mode="disabled")
    env = GroupedActionsObservations(base)
    env.reset()
    orcs = [OracleEnv(width=W, height=H, gravity=False, queue_size=3) for _ in range(n)]
    boards = np.empty((n, H + 4, W + 8), np.uint8)
    pieces, rots = rng.integers(0, 7, n), rng.integers(0, 4, n)
    for i, o in enumerate(orcs):
        o.set_sequence(seqs[i])
        o.reset()
        b = o.board
        k = int(rng.integers(1, 6))                       # full rows at the bottom ...
        b[H - k:H, 4:4 + W] = rng.integers(2, 9, size=(k, W))
        wells = rng.choice(W, size=int(rng.integers(1, 3)), replace=False)
        depth = int(rng.integers(1, k + 1))
        b[H - k:H - k + depth, 4 + wells] = 0            # ... but for wells of random depth
        noise = rng.random((3, W)) < 0.3                  # some rubble above
        b[H - k - 3:H - k, 4:4 + W] = np.where(noise, rng.integers(2, 9, size=(3, W)), 0)
        o.board = b
        o.set_active(int(pieces[i]), int(rots[i]))
        boards[i] = b
    base.set_state(board=boards, piece=pieces, rotation=rots)
    got = np_(env.observation())
    cleared = 0
    want_legal = np.empty((n, A), np.uint8)
    for i, o in enumerate(orcs):
        _, wb, wl = o.grouped_observe(features=False, boards=True)
        _, _, ln = o.grouped_observe_lines()
        cleared += int((ln > 0).sum())
        want_legal[i] = wl
        if not np.array_equal(got[i], wb):
            bad = np.flatnonzero((got[i] != wb).reshape(A, -1).any(1))
            raise AssertionError(f"env {i}: placement images {bad[:6]} differ\n got:\n{got[i][bad[0]]}\n want:\n{wb[bad[0]]}")
    assert np.array_equal(np_(env.legal_actions_mask), want_legal)
    assert cleared > n, "the constructed boards should produce many row-clearing placements"
    # execute one legal placement per env and compare the locked boards and the re-enumeration
    a = np.array([int(rng.choice(np.flatnonzero(want_legal[i]))) for i in range(n)])
    g, r, term, _, info = env.step(torch.from_numpy(a))
    for i, o in enumerate(orcs):
        code, rr, tt, ll = o.grouped_step(int(a[i]), True)
        assert (float(np_(r)[i]), bool(np_(term)[i]), int(np_(info["lines_cleared"])[i])) == (np.float32(rr), tt, ll), i
        _, wb, wl = o.grouped_observe(features=False, boards=True)
        assert np.array_equal(np_(g)[i], wb) and np.array_equal(np_(info["action_mask"])[i], wl), i
    base.close()


@pytest.mark.parametrize("W,H", [(10, 20), (20, 40), (10, 40), (20, 24), (13, 30)], ids=lambda v: str(v))
def test_grouped_features_on_constructed_boards_vs_oracle(W, H):
    """Feature enumeration (GroupedActionsObservations + FeatureVectorObservation) on constructed boards: nearly full bottom
    rows with wells (many of the 4W placements clear 1-4 rows), rubble, and for a third of the envs stacks that reach the
    spawn rows (placements that end in row 0 / game-over placements).  (10,20) (20,40) (10,40) (20,24) take the packed-byte
    kernel with u32 / u64 columns, (13,30) the generic one."""
    from gpu_util import np_
    from oracle.tetris_oracle import OracleEnv
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations

    n, A = 96, 4 * W
    rng = np.random.default_rng(11 * W + H)
    seqs = rng.integers(0, 7, size=(n, 31)).astype(np.uint8)
    base = Tetris(width=W, height=H, gravity=False, queue_size=3, num_envs=n, randomizer_mode="sequence",
                  piece_sequences=seqs, autoreset_mode="disabled")
    env = GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)])
    env.reset()
    orcs = [OracleEnv(width=W, height=H, gravity=False, queue_size=3) for _ in range(n)]
    boards = np.empty((n, H + 4, W + 8), np.uint8)
    pieces, rots = rng.integers(0, 7, n), rng.integers(0, 4, n)
    for i, o in enumerate(orcs):
        o.set_sequence(seqs[i])
        o.reset()
        b = o.board
        k = int(rng.integers(1, 6))
        b[H - k:H, 4:4 + W] = rng.integers(2, 9, size=(k, W))
        wells = rng.choice(W, size=int(rng.integers(1, 3)), replace=False)
        depth = int(rng.integers(1, k + 1))
        b[H - k:H - k + depth, 4 + wells] = 0
        noise = rng.random((3, W)) < 0.3
        b[H - k - 3:H - k, 4:4 + W] = np.where(noise, rng.integers(2, 9, size=(3, W)), 0)
        if i % 3 == 0:                                    # towers into the spawn rows, holes inside
            for c in rng.choice(W, size=int(rng.integers(1, W)), replace=False):
                top = int(rng.integers(0, 6))
                col = np.where(rng.random(H - k - top) < 0.8, rng.integers(2, 9, size=H - k - top), 0)
                col[0] = 3
                b[top:H - k, 4 + c] = col
        o.board = b
        o.set_active(int(pieces[i]), int(rots[i]))
        boards[i] = b
    base.set_state(board=boards, piece=pieces, rotation=rots)
    got = np_(env.observation())
    cleared = kinds2 = 0
    want_legal = np.empty((n, A), np.uint8)
    for i, o in enumerate(orcs):
        wf, _, wl = o.grouped_observe(features=True, boards=False)
        _, _, ln = o.grouped_observe_lines()
        cleared += int((ln > 0).sum())
        kinds2 += int(((wf == 0).all(1) & (wl > 0)).sum())
        want_legal[i] = wl
        if not np.array_equal(got[i], wf):
            bad = np.flatnonzero((got[i] != wf).any(1))
            raise AssertionError(f"env {i} piece {pieces[i]} rot {rots[i]}: placements {bad[:6]} differ\n got:\n{got[i][bad[:6]]}\n want:\n{wf[bad[:6]]}")
    assert np.array_equal(np_(env.legal_actions_mask), want_legal)
    assert cleared > n and kinds2 > 0
    a = np.array([int(rng.choice(np.flatnonzero(want_legal[i]))) for i in range(n)])
    g, r, term, _, info = env.step(torch.from_numpy(a))
    for i, o in enumerate(orcs):
        code, rr, tt, ll = o.grouped_step(int(a[i]), True)
        assert (float(np_(r)[i]), bool(np_(term)[i]), int(np_(info["lines_cleared"])[i])) == (np.float32(rr), tt, ll), i
        wf, _, wl = o.grouped_observe(features=True, boards=False)
        assert np.array_equal(np_(g)[i], wf) and np.array_equal(np_(info["action_mask"])[i], wl), i
        assert np.array_equal(np_(info["board"])[i], o.features(o.obs())), i
    base.close()


@pytest.mark.parametrize("n,terminate,autoreset", [(4096 + 17, True, "next_step"), (1000, False, "same_step"), (33, True, "disabled")])
def test_fused_grouped_step_equals_two_kernel_path(n, terminate, autoreset, monkeypatch):
    """tg_grouped_step has two implementations for the feature observation of the 10-wide board: one persistent kernel
    (k_grouped_step_feats: logic warps + feature warps, small batches) and k_step_ws<.., 2> followed by k_grouped_feats_x.  Same
    seeds, same actions (random legal placements with some illegal ones): every output of every step must be identical."""
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations

    def make():
        base = Tetris(num_envs=n, gravity=False, queue_size=4, autoreset_mode=autoreset)
        return base, GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)], terminate_on_illegal_action=terminate)

    (ba, ea), (bb, eb) = make(), make()
    monkeypatch.setenv("TG_GROUPED_SPLIT", "1")
    oa, ia = ea.reset(seed=7)
    ob, ib = eb.reset(seed=7)
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(60):
        a = torch.multinomial(ea.legal_actions_mask.float() + 1e-9, 1, generator=g).squeeze(1).to(torch.int32)
        if t % 7 == 3:   # some arbitrary (often illegal) placements
            a = torch.where(torch.rand(n, device="cuda", generator=g) < 0.2, torch.randint(0, 40, (n,), device="cuda", generator=g, dtype=torch.int32), a)
        monkeypatch.setenv("TG_GROUPED_SPLIT", "1")
        monkeypatch.delenv("TG_GROUPED_FUSED", raising=False)
        ra = ea.step(a)
        monkeypatch.delenv("TG_GROUPED_SPLIT")
        monkeypatch.setenv("TG_GROUPED_FUSED", "1")
        rb = eb.step(a)
        torch.cuda.synchronize()
        assert torch.equal(ra[0], rb[0]), f"features differ at step {t}"
        assert torch.equal(ra[1], rb[1]) and torch.equal(ra[2], rb[2]) and torch.equal(ra[3], rb[3]), f"5-tuple differs at step {t}"
        for k in ("action_mask", "board", "lines_cleared"):
            assert torch.equal(ra[4][k], rb[4][k]), f"info[{k}] differs at step {t}"
        for k in ("_hot", "_brd", "_rng"):   # the packed state records themselves
            assert torch.equal(getattr(ba, k), getattr(bb, k)), f"state {k} differs at step {t}"

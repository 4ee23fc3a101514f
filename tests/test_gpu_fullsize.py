"""Size-independent properties at BASELINE.json's full size (4,194,304 envs, 10x20, queue 7, obs dict every step), where the
C oracle is far too slow to follow: invariants of the game that tie the observation writer, the state records and the 5-tuple
together, determinism, and shard invariance (the first 65,536 envs of the big batch equal a 65,536-env run).  The small-batch
tests (test_gpu_base.py, test_gpu_10k_episodes.py) establish bit-exactness against the oracle; these check that nothing changes
with size (tile scheduling, persistent CTAs, 32-bit offsets: 4 M envs x 432 B images exceed 2^31 bytes per array)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

N, T, W, H, Q = 1 << 22, 48, 10, 20, 7


def _run(n, acts, offset=0):
    from tetris_gymnasium_b200.envs.tetris import Tetris

    env = Tetris(num_envs=n, queue_size=Q, autoreset_mode="next_step", env_id_offset=offset)
    obs, _ = env.reset(seed=42)
    lines_acc = torch.zeros(n, dtype=torch.int32, device="cuda")
    prev_term = torch.zeros(n, dtype=torch.bool, device="cuda")
    total_term = 0
    for t in range(acts.shape[0]):
        obs, r, term, trunc, info = env.step(acts[t, :n])
        lines_acc = torch.where(prev_term, torch.zeros_like(lines_acc), lines_acc) + info["lines_cleared"]
        prev_term = term.clone()
        total_term += int(term.sum())
        assert not bool(trunc.any())
    return env, obs, lines_acc, total_term


def test_full_size_invariants_determinism_and_shard_invariance():
    if torch.cuda.mem_get_info()[0] < 40 * 2**30:
        pytest.skip("needs about 40 GB of free device memory")
    g = torch.Generator(device="cuda")
    g.manual_seed(42)
    acts = torch.randint(0, 8, (T, N), dtype=torch.int32, device="cuda", generator=g)
    acts[4::5] = 5                                                # every fifth step: hard drop for everyone (commits, line clears)
    env, obs, lines_acc, total_term = _run(N, acts)
    st = env.get_state()
    board = st["board"]                                           # locked cells, u8 [N, 24, 18]
    field = board[:, :H, 4:4 + W]
    # bedrock frame intact, no filled row survives a commit
    assert bool((board[:, H:, :] == 1).all()) and bool((board[:, :H, :4] == 1).all()) and bool((board[:, :H, 4 + W:] == 1).all())
    assert not bool((field != 0).all(dim=2).any())
    assert int(field.max()) <= 8
    # every commit adds 4 cells, every cleared row removes W: cells + W * lines == 0 (mod 4) since the env's last reset
    cells = (field != 0).flatten(1).sum(1).to(torch.int32)
    assert bool(((cells + W * lines_acc) % 4 == 0).all())
    # observation dict vs state: board image = locked cells + the <= 4 cells of the active piece, inside the n x n mask box
    diff = obs["board"] != board
    nd = diff.flatten(1).sum(1)
    assert bool(((nd == 0) | (nd == 4)).all())
    assert not bool((diff & (obs["active_tetromino_mask"] == 0)).any())
    piece = st["piece"]
    nbox = torch.where(piece == 0, 4, torch.where(piece == 1, 2, 3))
    assert torch.equal(obs["active_tetromino_mask"].flatten(1).sum(1).to(torch.int64), (nbox * nbox).to(torch.int64))
    vals = torch.where(diff, obs["board"], torch.zeros_like(board)).flatten(1).max(1).values
    assert bool(((nd == 0) | (vals.to(torch.int32) == piece + 2)).all())
    # queue / holder images carry the ids of the state's queue and holder
    qimg = obs["queue"].view(N, 4, Q, 4)
    assert torch.equal(qimg.amax(dim=(1, 3)).to(torch.int32), st["queue"] + 2)
    assert bool(((qimg != 0).sum(dim=(1, 3)) == 4).all())
    hp = st["holder_piece"]
    hmax = obs["holder"].flatten(1).max(1).values.to(torch.int32)
    assert bool(torch.where(hp < 0, hmax == 1, hmax == hp + 2).all())
    assert total_term > 0 and int(lines_acc.sum()) > 0
    # determinism: the same run again gives the same bytes
    keep_board, keep_obs = board.clone(), obs["board"].clone()
    keep_x, keep_queue = st["x"].clone(), st["queue"].clone()
    env.close()
    del env, obs, st, board, field, diff
    torch.cuda.empty_cache()
    env2, obs2, _, total_term2 = _run(N, acts)
    st2 = env2.get_state()
    assert total_term2 == total_term
    assert torch.equal(st2["board"], keep_board) and torch.equal(obs2["board"], keep_obs)
    assert torch.equal(st2["x"], keep_x) and torch.equal(st2["queue"], keep_queue)
    env2.close()
    del env2, obs2, st2
    torch.cuda.empty_cache()
    # shard invariance: Philox streams are keyed by the global env id, so a 65,536-env run reproduces the first / a middle
    # slice of the big batch (env_id_offset = global id of local env 0; per-env seeds = 42 + global id)
    for off in (0, 3 * 65536 + 32):
        from tetris_gymnasium_b200.envs.tetris import Tetris

        m = 65536
        env3 = Tetris(num_envs=m, queue_size=Q, autoreset_mode="next_step", env_id_offset=off)
        env3.reset(seed=torch.arange(off, off + m, dtype=torch.int64) + 42)
        for t in range(T):
            o3, *_ = env3.step(acts[t, off:off + m].contiguous())
        s3 = env3.get_state()
        assert torch.equal(s3["board"], keep_board[off:off + m]) and torch.equal(o3["board"], keep_obs[off:off + m])
        assert torch.equal(s3["queue"], keep_queue[off:off + m])
        env3.close()


def test_full_size_grouped_feature_rows_are_self_consistent():
    """1,048,576 envs x 40 placements (BASELINE config 3 at its largest size): every feature row must satisfy the identities
    of wrappers/observation.py:238-278 -- max height = max of the heights, bumpiness = sum |h[c+1] - h[c]| (uint8 wrap) -- the
    frame / game-over patterns of wrappers/grouped.py:160-170 must agree with the legal mask, and info["board"] must equal the
    feature wrapper applied to the real observation (tg_features, an independent kernel)."""
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations

    n, A, F = 1 << 20, 4 * W, W + 3
    base = Tetris(num_envs=n, gravity=False, queue_size=4)
    env = GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)])
    feats, info = env.reset(seed=7)
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    ref_feats = torch.empty((n, F), dtype=torch.uint8, device="cuda")
    for t in range(24):
        legal = info["action_mask"]
        h = feats[:, :, :W].to(torch.int32)
        assert torch.equal(feats[:, :, W].to(torch.int32), h.amax(dim=2)), t
        bump = (h[:, :, 1:] - h[:, :, :-1]).abs().sum(dim=2) & 255
        assert torch.equal(feats[:, :, W + 2].to(torch.int32), bump), t
        assert int(h.max()) <= H
        # illegal (frame) placements: ones board with row 0 zeroed -> heights and max H - 1, no holes, no bumpiness
        ill = legal == 0
        want = torch.tensor([H - 1] * (W + 1) + [0, 0], dtype=torch.uint8, device="cuda")
        assert bool((feats[ill] == want).all()), t
        assert bool(legal.any(dim=1).all())
        # info["board"] vs the stand-alone feature kernel on the same state
        from tetris_gymnasium_b200 import _lib
        _lib.check(base._L.tg_features(base._h, base._state(), n, ref_feats.data_ptr(), base._stream()), base._h)
        assert torch.equal(info["board"], ref_feats), t
        a = torch.multinomial(legal.float() + 1e-9, 1, generator=g).squeeze(1).to(torch.int32)
        feats, r, term, trunc, info = env.step(a)
    base.close()


def test_full_size_wide_board_rgb_equals_palette_of_the_obs_dict():
    """BASELINE config 5 (wide 20x40 board, queue 5, RGB image) at 262,144 envs: the image written by k_rgb must be the palette
    lookup of the observation dict written by the step kernel for the same state -- Tetris.get_rgb (envs/tetris.py:309-343)
    restated with torch ops: board | (queue, ones padding, holder) -> colours."""
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import RgbObservation

    n, Wd, Hd, Qd = 1 << 18, 20, 40, 5
    base = Tetris(num_envs=n, width=Wd, height=Hd, queue_size=Qd)
    env = RgbObservation(base, keep_obs_dict=True)
    env.reset(seed=11)
    colors = torch.tensor([[0, 0, 0], [128, 128, 128], [0, 240, 240], [240, 240, 0], [160, 0, 240], [0, 240, 0], [240, 0, 0],
                           [0, 0, 240], [240, 160, 0]], dtype=torch.uint8, device="cuda")      # envs/tetris.py:45-75
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    for t in range(12):
        a = torch.randint(0, 8, (n,), dtype=torch.int32, device="cuda", generator=g)
        if t % 3 == 2:
            a.fill_(5)
        img, *_ = env.step(a)
        obs = base._obs()
        Hp, max_len = Hd + 4, 4 * Qd
        holder = torch.cat([obs["holder"], torch.ones((n, 4, max_len - 4), dtype=torch.uint8, device="cuda")], dim=2)
        pad = torch.ones((n, Hp - 8, max_len), dtype=torch.uint8, device="cuda")
        stack = torch.cat([obs["board"], torch.cat([obs["queue"], pad, holder], dim=1)], dim=2)
        assert img.shape == (n, Hp, Wd + 8 + max_len, 3)
        assert torch.equal(img, colors[stack.long()]), t
    base.close()

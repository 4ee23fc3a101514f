"""GPU parity of the functional facade (tg_fn_step through envs/tetris_fn.py) against the numpy restatement
of the reference functional env, with injected bags."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H,gravity", [(10, 20, True), (10, 20, False), (8, 12, True), (20, 40, True)])
def test_functional_facade_vs_oracle(W, H, gravity):
    from oracle.tetris_fn_oracle import FnOracle
    from tetris_gymnasium_b200.envs import tetris_fn as fn
    from tetris_gymnasium_b200.functional import TETROMINOES, EnvConfig

    B, T, Q = 96, 300, 7
    rng = np.random.default_rng(W * H)
    seqs = np.stack([np.concatenate([rng.permutation(7) for _ in range(9)]) for _ in range(B)]).astype(np.uint8)
    cfg = EnvConfig(width=W, height=H, padding=4, queue_size=Q, gravity_enabled=gravity)
    keys = torch.zeros((B, 2), dtype=torch.int64)
    keys, states, obs = fn.batched_reset(TETROMINOES, keys, config=cfg, create_queue_fn=seqs)
    orcs = [FnOracle(W, H, Q, gravity, seq=seqs[i]) for i in range(B)]
    want = np.stack([o.reset() for o in orcs])
    assert obs.dtype == torch.int8 and np.array_equal(obs.cpu().numpy(), want)
    n_over = 0
    for t in range(T):
        a = rng.integers(0, 7, size=B)
        states, obs, reward, term, info = fn.batched_step(TETROMINOES, states, a, config=cfg, queue_fn=seqs)
        res = [o.step(int(a[i])) for i, o in enumerate(orcs)]
        assert np.array_equal(obs.cpu().numpy(), np.stack([r[0] for r in res])), t
        assert np.array_equal(reward.cpu().numpy(), np.array([r[1] for r in res], np.float32)), t
        assert np.array_equal(term.cpu().numpy(), np.array([r[2] for r in res])), t
        assert np.array_equal(info["lines_cleared"].cpu().numpy(), np.array([r[3] for r in res])), t
        n_over = int(term.sum())
    assert np.array_equal(states.board.cpu().numpy(), np.stack([o.board for o in orcs]))
    assert np.array_equal(states.score.cpu().numpy(), np.array([o.score for o in orcs], np.float32))
    assert n_over > 0


def test_functional_single_env_signature_and_philox_bags():
    from tetris_gymnasium_b200.envs import tetris_fn as fn
    from tetris_gymnasium_b200.functional import TETROMINOES, EnvConfig

    cfg = EnvConfig(width=10, height=20, padding=4, queue_size=7)
    key, state, obs = fn.reset(TETROMINOES, torch.tensor([0, 42]), cfg)
    assert obs.shape == (20, 10) and sorted(state.queue[0].tolist()) == list(range(7))
    state2, obs2, reward, terminated, info = fn.step(TETROMINOES, state, 5, cfg)
    assert float(reward) == float(state2.score[0]) - float(state.score[0]) and "lines_cleared" in info
    # a finished game is frozen (test_env/test_step.py:16-25)
    over = state.replace(game_over=torch.ones_like(state.game_over))
    s3, _, r3, t3, _ = fn.step(TETROMINOES, over, 0, cfg)
    assert bool(t3) and float(r3) == 0.0 and torch.equal(s3.board, over.board)
    # every refill is a permutation
    st = state
    seen = []
    for _ in range(40):
        st, _, _, term, _ = fn.step(TETROMINOES, st, 6, cfg)
        seen.append(sorted(st.queue[0].tolist()))
    assert all(s == list(range(7)) for s in seen)


def test_functional_state_record_cache_is_invalidated_by_edits():
    """`_pack` reuses the record array a State came with only while its fields are untouched: attribute assignment,
    in-place edits and replace() must all be seen by the next step."""
    from tetris_gymnasium_b200.envs import tetris_fn as fn
    from tetris_gymnasium_b200.functional import TETROMINOES, EnvConfig
    from tetris_gymnasium_b200.functional.core import State

    cfg = EnvConfig(width=10, height=20, padding=4, queue_size=7)
    B = 64
    keys = torch.stack([torch.arange(B), torch.arange(B) + 7], dim=1)
    _, st, _ = fn.batched_reset(TETROMINOES, keys, config=cfg)
    a = torch.full((B,), 2, dtype=torch.int32, device="cuda")

    def fresh(s):   # a State without the cached record array (forces the slow packing path)
        return State(**{k: getattr(s, k).clone() for k in ("rng_key", "board", "active_tetromino", "rotation", "x", "y", "queue",
                                                            "queue_index", "game_over", "score")})

    def same(s1, s2):
        return all(torch.equal(getattr(s1, k), getattr(s2, k)) for k in ("board", "x", "y", "rotation", "active_tetromino", "queue", "score"))

    for _ in range(5):                                   # untouched loop: cached record array
        ref, *_ = fn.batched_step(TETROMINOES, fresh(st), a, config=cfg)
        st, *_ = fn.batched_step(TETROMINOES, st, a, config=cfg)
        assert same(st, ref)
    st.x.add_(1)                                         # in-place edit of a view field
    ref, *_ = fn.batched_step(TETROMINOES, fresh(st), a, config=cfg)
    nxt, *_ = fn.batched_step(TETROMINOES, st, a, config=cfg)
    assert same(nxt, ref)
    st = nxt
    st.score += 5.0                                      # in-place edit of a copied field
    ref, *_ = fn.batched_step(TETROMINOES, fresh(st), a, config=cfg)
    nxt, *_ = fn.batched_step(TETROMINOES, st, a, config=cfg)
    assert same(nxt, ref) and float(nxt.score.min()) >= 5.0
    st = nxt
    st.rotation = (st.rotation + 1) & 3                  # attribute assignment
    ref, *_ = fn.batched_step(TETROMINOES, fresh(st), a, config=cfg)
    nxt, *_ = fn.batched_step(TETROMINOES, st, a, config=cfg)
    assert same(nxt, ref)
    st = nxt.replace(y=nxt.y + 1)                        # replace(): a new State without the tag
    ref, *_ = fn.batched_step(TETROMINOES, fresh(st), a, config=cfg)
    nxt, *_ = fn.batched_step(TETROMINOES, st, a, config=cfg)
    assert same(nxt, ref)


def test_functional_uniform_queue():
    """queue.create_uniform_queue (functional/queue.py:71-119): every refill draws queue_size values from [0, queue_size - 1)
    (randint's maxval is exclusive: the reference never yields the last piece), index restarts at 1."""
    from tetris_gymnasium_b200.envs import tetris_fn as fn
    from tetris_gymnasium_b200.functional import TETROMINOES, EnvConfig, create_uniform_queue, uniform_queue_get_next_element

    cfg = EnvConfig(width=10, height=20, padding=4, queue_size=7)
    B = 512
    keys = torch.stack([torch.arange(B), torch.arange(B) * 3 + 1], dim=1)
    _, st, _ = fn.batched_reset(TETROMINOES, keys, config=cfg, create_queue_fn=create_uniform_queue)
    q0 = st.queue.clone()
    assert int(q0.min()) >= 0 and int(q0.max()) == 5 and bool((st.queue_index == 1).all())
    assert torch.equal(st.active_tetromino, q0[:, 0])
    assert len({tuple(r) for r in q0.tolist()}) > B // 2        # not permutations, env-dependent
    assert any(len(set(r)) < 7 for r in q0.tolist())            # repeats occur (a bag would never repeat)
    a = torch.full((B,), 6, dtype=torch.int32, device="cuda")   # hard drops: one piece per step
    seen = [q0]
    for t in range(30):
        st, _, _, term, _ = fn.batched_step(TETROMINOES, st, a, config=cfg, queue_fn=uniform_queue_get_next_element)
        seen.append(st.queue.clone())
    allq = torch.stack(seen)
    assert int(allq.max()) == 5 and int(allq.min()) == 0
    assert not torch.equal(seen[0], seen[-1])                    # refilled along the way
    # same key, same selector -> same stream; the bag selector gives permutations instead
    _, st2, _ = fn.batched_reset(TETROMINOES, keys, config=cfg, create_queue_fn="uniform")
    assert torch.equal(st2.queue, q0)
    _, st3, _ = fn.batched_reset(TETROMINOES, keys, config=cfg)
    assert all(sorted(r) == list(range(7)) for r in st3.queue.tolist())

"""Host logic of the functional facade that needs no GPU: the State <-> record-array packing and the cache that lets the
usual `state = step(state)` loop skip the packing (tetris_gymnasium_b200/envs/tetris_fn.py)."""
import torch

from tetris_gymnasium_b200.envs import tetris_fn as fn
from tetris_gymnasium_b200.functional import EnvConfig, create_bag_queue, create_uniform_queue, uniform_queue_get_next_element
from tetris_gymnasium_b200.functional.core import State


def _records(B=5, Q=7):
    g = torch.Generator().manual_seed(1)
    sc = torch.randint(0, 7, (B, fn._S + Q), dtype=torch.int32, generator=g)
    sc[:, 6] = torch.tensor([0.0, 1.0, 100.0, 800.0, 3.0]).view(torch.int32)      # score as float32 bits
    sc[:, 7] = torch.tensor([0, 1, -1, 2**31 - 1, -2**31], dtype=torch.int32)     # u32 keys with the high bit set
    sc[:, 5] = torch.tensor([0, 1, 0, 0, 1], dtype=torch.int32)
    board = torch.zeros((B, 24, 18), dtype=torch.int8)
    return board, sc


def test_unpack_pack_round_trip_and_views():
    board, sc = _records()
    st = fn._unpack(board, sc)
    assert st.rng_key.dtype == torch.int64 and int(st.rng_key.min()) >= 0 and int(st.rng_key[2, 0]) == 2**32 - 1
    assert st.score.dtype == torch.float32 and st.score.tolist() == [0.0, 1.0, 100.0, 800.0, 3.0]
    assert st.game_over.dtype == torch.bool and st.game_over.tolist() == [False, True, False, False, True]
    assert fn._pack(st, 7) is sc                                   # untouched: the record array itself
    fresh = State(**{k: getattr(st, k).clone() for k in ("rng_key", "board", "active_tetromino", "rotation", "x", "y", "queue",
                                                         "queue_index", "game_over", "score")})
    assert torch.equal(fn._pack(fresh, 7), sc)                    # slow path reproduces it bit for bit


def test_cache_is_dropped_on_any_edit():
    board, sc = _records()
    st = fn._unpack(board, sc)
    st.x.add_(1)                                                   # in-place edit of a view field (aliases the records)
    p = fn._pack(st, 7)
    assert p is not sc and torch.equal(p[:, 2], st.x)
    st = fn._unpack(board, sc.clone())
    st.score += 5.0                                                # in-place edit of a copied field
    assert fn._pack(st, 7)[:, 6].view(torch.float32).tolist() == st.score.tolist()
    st = fn._unpack(board, sc.clone())
    st.rotation = (st.rotation + 1) & 3                            # attribute assignment
    assert torch.equal(fn._pack(st, 7)[:, 1], st.rotation)
    st2 = fn._unpack(board, sc.clone()).replace(y=torch.full((5,), 9, dtype=torch.int32))   # replace(): no tag
    assert not hasattr(st2, "_tg_packed") and fn._pack(st2, 7)[:, 3].tolist() == [9] * 5


def test_queue_selectors():
    assert fn._seq(None, "cpu") is None and fn._seq("bag", "cpu") is None and fn._seq(create_bag_queue, "cpu") is None
    assert fn._seq("uniform", "cpu") == fn.UNIFORM and fn._seq(create_uniform_queue, "cpu") == fn.UNIFORM
    assert fn._seq(uniform_queue_get_next_element, "cpu") == fn.UNIFORM
    seq = fn._seq([[0, 1, 2], [3, 4, 5]], "cpu")
    assert seq.dtype == torch.uint8 and seq.shape == (2, 3)
    # called directly they need the device (they run the facade's own bag routine): without one they fail loudly, no CPU path
    if not torch.cuda.is_available():
        try:
            create_uniform_queue(EnvConfig(width=10, height=20, padding=4, queue_size=7), torch.tensor([0, 1]))
            raise AssertionError("no CPU fallback expected")
        except RuntimeError:
            pass

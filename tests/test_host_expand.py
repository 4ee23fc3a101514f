"""CPU: the host-side dict expansion of the compact host step (tg_host_expand, csrc/tg_host_expand*) against the oracle.

The expansion is a format conversion: packed records (hot word + nibble id plane, the layout of DESIGN.md section 2) -> the
observation dict of Tetris._get_obs (envs/tetris.py:566-615).  Here the records are packed in Python from ORACLE states, so
the test needs no GPU; tests/test_gpu_host_step.py compares the same code with the device-written dict on real records."""
import ctypes as C

import numpy as np
import pytest

from oracle.tetris_oracle import OracleEnv
from tetris_gymnasium_b200 import _lib

BASE = [np.array(m, np.uint8) for m in (
    [[0, 0, 0, 0], [1, 1, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]], [[1, 1], [1, 1]], [[0, 1, 0], [1, 1, 1], [0, 0, 0]],
    [[0, 1, 1], [1, 1, 0], [0, 0, 0]], [[1, 1, 0], [0, 1, 1], [0, 0, 0]], [[1, 0, 0], [1, 1, 1], [0, 0, 0]],
    [[0, 0, 1], [1, 1, 1], [0, 0, 0]])]


def layout(W, H):
    Hp = H + 4
    ids_off = W * (8 if Hp > 32 else 4)
    ids_words = (H * W + 7) // 8
    bs = (ids_off + 4 * ids_words + 15) // 16 * 16
    if (bs // 16) % 2 == 0:
        bs += 16
    return ids_off, bs


def rotation_of(piece, matrix):
    for r in range(4):
        if np.array_equal(np.rot90(BASE[piece], k=r) > 0, matrix > 0):
            return r
    raise AssertionError("matrix is no rotation of its base")


def pack_records(envs):
    """hot u8[n][32] and board u8[n][board_stride] (id plane only; the expansion never reads the column bitboards)."""
    e0 = envs[0]
    W, H = e0.W, e0.H
    ids_off, bs = layout(W, H)
    hot = np.zeros((len(envs), 8), np.uint32)
    brd = np.zeros((len(envs), bs), np.uint8)
    for i, e in enumerate(envs):
        s = e.scalars()
        p = s["active"]
        r = rotation_of(p, e.active_matrix())
        hold, hold_r = 0, 0
        if e.holder_size > 1:      # FIFO word: count | slots (piece | rotation << 3), oldest first
            hq = 0
            for k, (idx, m) in enumerate(e.held_slots()):
                hq |= (idx | (rotation_of(idx, m) << 3)) << (3 + 5 * k)
            hot[i, 7] = hq | e.holder_len()
        else:
            hm = e.held_matrix()
            if hm is not None:
                hold = s["holder"] + 1
                hold_r = rotation_of(s["holder"], hm)
        hot[i, 0] = s["x"] | (s["y"] << 6) | (p << 13) | (r << 16) | (hold << 18) | (hold_r << 22)
        q = 0
        for k, v in enumerate(s["queue"]):
            q |= int(v) << (4 * k)
        hot[i, 2], hot[i, 3] = q & 0xFFFFFFFF, q >> 32
        cells = e.board[:H, 4:4 + W].reshape(-1).astype(np.uint8)
        if len(cells) % 2:
            cells = np.append(cells, 0)
        brd[i, ids_off:ids_off + len(cells) // 2] = cells[0::2] | (cells[1::2] << 4)
    return hot.view(np.uint8).reshape(len(envs), 32), brd


def expand(W, H, Q, hot, brd, threads=2, misalign=0, holder_size=1):
    n = len(hot)
    L = _lib.load()
    cfg = _lib.TgConfig()
    cfg.width, cfg.height, cfg.queue_size, cfg.holder_size = W, H, Q, holder_size
    for i in range(8):
        cfg.action_map[i] = i
    Hp, Wp = H + 4, W + 8

    def buf(shape):
        raw = np.full(int(np.prod(shape)) + 128, 0xEE, np.uint8)
        off = (-raw.ctypes.data) % 64 + misalign
        return raw[off:off + int(np.prod(shape))].reshape(shape)

    out = {"board": buf((n, Hp, Wp)), "active_tetromino_mask": buf((n, Hp, Wp)), "holder": buf((n, 4, 4 * holder_size)), "queue": buf((n, 4, 4 * Q))}
    ho = _lib.TgObs(out["board"].ctypes.data, out["active_tetromino_mask"].ctypes.data, out["holder"].ctypes.data, out["queue"].ctypes.data)
    hot = np.ascontiguousarray(hot)
    brd = np.ascontiguousarray(brd)
    rc = L.tg_host_expand(C.byref(cfg), n, hot.ctypes.data, brd.ctypes.data, ho, threads)
    assert rc == 0, L.tg_last_error(None)
    return out


@pytest.mark.parametrize("W,H,Q,n,steps,misalign", [(10, 20, 7, 300, 40, 0), (10, 20, 4, 131, 25, 0), (20, 40, 5, 70, 60, 0),
                                                    (7, 9, 3, 150, 30, 0), (13, 21, 1, 97, 30, 0), (24, 12, 16, 40, 20, 0),
                                                    (10, 20, 7, 130, 10, 16), (10, 8, 2, 260, 12, 0)])
def test_expand_matches_oracle_obs(W, H, Q, n, steps, misalign):
    rng = np.random.default_rng(W * 100 + H + Q)
    envs = [OracleEnv(width=W, height=H, queue_size=Q, gravity=True) for _ in range(n)]
    for i, e in enumerate(envs):
        e.set_sequence(rng.integers(0, 7, size=64).astype(np.uint8))
        e.reset()
    for t in range(steps):
        for e in envs:
            if not e.scalars()["game_over"]:
                e.step(int(rng.integers(0, 8)))
        if t % 5 == 4 or t == steps - 1:
            hot, brd = pack_records(envs)
            got = expand(W, H, Q, hot, brd, threads=1 + t % 3, misalign=misalign)
            want = [e.obs() for e in envs]
            for k in got:
                w = np.stack([o[k] for o in want])
                assert np.array_equal(got[k], w), (k, t, np.flatnonzero((got[k] != w).reshape(n, -1).any(1))[:5])


@pytest.mark.parametrize("S,n", [(2, 200), (3, 131), (4, 64)])
def test_expand_with_a_fifo_holder(S, n):
    rng = np.random.default_rng(S)
    envs = [OracleEnv(queue_size=4, holder_size=S) for _ in range(n)]
    for e in envs:
        e.set_sequence(rng.integers(0, 7, size=64).astype(np.uint8))
        e.reset()
    for t in range(60):
        for e in envs:
            if not e.scalars()["game_over"]:
                e.step(int(rng.choice([0, 1, 3, 5, 6, 6, 6, 7])))
        if t % 6 == 5:
            hot, brd = pack_records(envs)
            got = expand(10, 20, 4, hot, brd, holder_size=S)
            want = [e.obs() for e in envs]
            for k in got:
                assert np.array_equal(got[k], np.stack([o[k] for o in want])), (k, t)


def test_expand_game_over_piece_not_projected():
    """A piece that collides where it stands (spawn blocked) is not projected into the board image (envs/tetris.py:566-576 goes
    through project_tetromino only for the non-colliding case in our kernels' reading: the oracle is the judge)."""
    e = OracleEnv(gravity=False)
    e.set_sequence(np.array([1] * 64, np.uint8))     # O pieces, hard drops in the spawn column until the stack blocks the spawn
    e.reset()
    for _ in range(40):
        if e.scalars()["game_over"]:
            break
        e.step(5)
    assert e.scalars()["game_over"]
    hot, brd = pack_records([e])
    got = expand(10, 20, 4, hot, brd)
    want = e.obs()
    for k in got:
        assert np.array_equal(got[k][0], want[k]), k

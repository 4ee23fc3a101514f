"""ctypes front-end of the C oracle (oracle/tetris_oracle.c) -- TEST INFRASTRUCTURE ONLY.

`OracleEnv` mirrors the shape of the reference API (reset/step returning the observation
dict; grouped_observe/grouped_step; features; rgb) so parity tests read like the reference's.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None


class OrcConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("gravity", C.c_int32),
        ("queue_size", C.c_int32),
        ("act", C.c_int32 * 8),
        ("r_alife", C.c_double),
        ("r_clear_line", C.c_double),
        ("r_game_over", C.c_double),
        ("r_invalid", C.c_double),
    ]


def lib():
    global _LIB
    if _LIB is None:
        path = _build.build()
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcConfig)]
        for name in (
            "orc_destroy orc_set_sequence orc_seed_numpy orc_get_obs orc_reset orc_features "
            "orc_grouped_observe orc_rgb orc_get_board orc_set_board orc_get_scalars "
            "orc_get_active_matrix orc_get_held_matrix orc_set_active orc_set_flags orc_set_holder "
            "orc_set_queue orc_vec_step orc_vec_grouped_step orc_rnd_stream orc_grouped_observe_ex orc_set_true_randomizer "
            "orc_vec_create orc_vec_destroy orc_vec_seed_words orc_vec_reset orc_set_holder_size orc_set_tetrominoes"
        ).split():
            getattr(L, name).restype = None
        L.orc_step.restype = C.c_int
        L.orc_grouped_step.restype = C.c_int
        L.orc_num_threads.restype = C.c_int
        L.orc_get_holder_len.restype = C.c_int
        L.orc_get_held_slot.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def numpy_pcg64_state(seed):
    """(state_hi, state_lo, inc_hi, inc_lo) of PCG64(SeedSequence(seed)) -- what
    Randomizer.reset builds for `seed > 0` (components/tetromino_randomizer.py:40-43)."""
    st = np.random.PCG64(np.random.SeedSequence(seed)).state["state"]
    m = (1 << 64) - 1
    return np.array([st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m], dtype=np.uint64)


class OracleEnv:
    """One reference-equivalent env.  Defaults follow envs/tetris.py:77-91 and the mappings."""

    ACTIONS = dict(move_left=0, move_right=1, move_down=2, rotate_clockwise=3,
                   rotate_counterclockwise=4, hard_drop=5, swap=6, no_op=7)

    def __init__(self, width=10, height=20, gravity=True, queue_size=4, actions=None,
                 alife=1.0, clear_line=1.0, game_over=0.0, invalid_action=-0.1, holder_size=1):
        L = lib()
        a = dict(self.ACTIONS)
        if actions:
            a.update(actions)
        cfg = OrcConfig()
        cfg.width, cfg.height, cfg.gravity, cfg.queue_size = width, height, int(bool(gravity)), queue_size
        order = ["move_left", "move_right", "move_down", "rotate_clockwise",
                 "rotate_counterclockwise", "hard_drop", "swap", "no_op"]
        for i, k in enumerate(order):
            cfg.act[i] = a[k]
        cfg.r_alife, cfg.r_clear_line, cfg.r_game_over, cfg.r_invalid = alife, clear_line, game_over, invalid_action
        self.cfg = cfg
        self.h = C.c_void_p(L.orc_create(C.byref(cfg)))
        if not self.h:
            raise ValueError("bad oracle config")
        self.holder_size = holder_size
        if holder_size != 1:
            L.orc_set_holder_size(self.h, int(holder_size))
        self.W, self.H, self.Q = width, height, queue_size
        self.Wp, self.Hp = width + 8, height + 4
        self._seq = None
        self.legal = np.ones(4 * width, dtype=np.uint8)

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass

    def set_tetrominoes(self, matrices, colors):
        """Tetris(tetrominoes=[...]): square binary matrices (<= 4 x 4) and RGB colours; call before reset."""
        n = np.array([len(m) for m in matrices], np.int32)
        mm = np.zeros((len(matrices), 16), np.uint8)
        for i, m in enumerate(matrices):
            mm[i, : n[i] * n[i]] = (np.asarray(m) != 0).astype(np.uint8).reshape(-1)
        rgb = np.ascontiguousarray(colors, dtype=np.uint8)
        lib().orc_set_tetrominoes(self.h, len(matrices), _p(n), _p(mm), _p(rgb))
        self._pieces = [np.asarray(m) != 0 for m in matrices]

    # -- randomizer ---------------------------------------------------------------------
    def set_sequence(self, seq, cursor=0):
        self._seq = np.ascontiguousarray(seq, dtype=np.uint8)
        lib().orc_set_sequence(self.h, _p(self._seq), C.c_int64(len(self._seq)), C.c_int64(cursor))

    def set_true_randomizer(self, on=True):
        """TrueRandomizer (components/tetromino_randomizer.py:105-136) instead of the 7-bag; call before seed_numpy."""
        lib().orc_set_true_randomizer(self.h, int(on))

    def seed_numpy(self, seed):
        st = numpy_pcg64_state(seed)
        lib().orc_seed_numpy(self.h, _p(st))

    def rnd_stream(self, n):
        """randomizer.reset() + n draws."""
        out = np.empty(n, np.uint8)
        lib().orc_rnd_stream(self.h, int(n), _p(out))
        return out

    # -- env ----------------------------------------------------------------------------
    def obs(self):
        o = {
            "board": np.empty((self.Hp, self.Wp), np.uint8),
            "active_tetromino_mask": np.empty((self.Hp, self.Wp), np.uint8),
            "holder": np.empty((4, 4 * self.holder_size), np.uint8),
            "queue": np.empty((4, 4 * self.Q), np.uint8),
        }
        lib().orc_get_obs(self.h, _p(o["board"]), _p(o["active_tetromino_mask"]), _p(o["holder"]), _p(o["queue"]))
        return o

    def reset(self, seed=None):
        if seed and seed > 0:
            self.seed_numpy(seed)
        lib().orc_reset(self.h)
        return self.obs(), {"lines_cleared": 0}

    def step(self, action):
        r, t, l = C.c_double(), C.c_int(), C.c_int()
        rc = lib().orc_step(self.h, int(action), C.byref(r), C.byref(t), C.byref(l))
        if rc != 0:
            raise AssertionError(f"{action!r} invalid")
        return self.obs(), r.value, bool(t.value), False, {"lines_cleared": l.value}

    # -- wrappers -----------------------------------------------------------------------
    def features(self, obs):
        """FeatureVectorObservation.observation on an obs dict (mutates obs['board'] like the reference)."""
        out = np.empty(self.W + 3, np.uint8)
        b = obs["board"]
        assert b.flags.c_contiguous and b.dtype == np.uint8
        m = np.ascontiguousarray(obs["active_tetromino_mask"], dtype=np.uint8)
        lib().orc_features(self.h, _p(b), _p(m), _p(out))
        return out

    def grouped_observe(self, features=True, boards=False):
        A = 4 * self.W
        f = np.empty((A, self.W + 3), np.uint8) if features else None
        b = np.empty((A, self.Hp, self.Wp), np.uint8) if boards else None
        lib().orc_grouped_observe(self.h, _p(b), _p(f), _p(self.legal))
        return f, b, self.legal.copy()

    def grouped_observe_lines(self):
        """(features u8[A, W+3], legal u8[A], lines i32[A]) -- lines: rows cleared, -1 game-over placement, -2 illegal."""
        A = 4 * self.W
        f = np.empty((A, self.W + 3), np.uint8)
        ln = np.empty(A, np.int32)
        lib().orc_grouped_observe_ex(self.h, None, _p(f), _p(self.legal), _p(ln))
        return f, self.legal.copy(), ln

    def grouped_step(self, action, terminate_on_illegal=True):
        r, t, l = C.c_double(), C.c_int(), C.c_int()
        code = lib().orc_grouped_step(self.h, int(action), _p(self.legal), int(terminate_on_illegal),
                                      C.byref(r), C.byref(t), C.byref(l))
        return code, r.value, bool(t.value), l.value

    def holder_len(self):
        return int(lib().orc_get_holder_len(self.h))

    def held_slots(self):
        """[(piece index, n x n id-valued matrix)] of the held pieces, oldest first."""
        out = []
        for s in range(self.holder_len()):
            n = C.c_int32()
            m = np.zeros(16, np.uint8)
            idx = lib().orc_get_held_slot(self.h, s, C.byref(n), _p(m))
            out.append((int(idx), m[: n.value * n.value].reshape(n.value, n.value).copy()))
        return out

    def rgb(self):
        out = np.empty((self.Hp, self.Wp + 4 * max(self.Q, self.holder_size, 1), 3), np.uint8)
        lib().orc_rgb(self.h, _p(out))
        return out

    # -- state poking (what the reference tests do through env.unwrapped.*) ----------------
    @property
    def board(self):
        b = np.empty((self.Hp, self.Wp), np.uint8)
        lib().orc_get_board(self.h, _p(b))
        return b

    @board.setter
    def board(self, b):
        b = np.ascontiguousarray(b, dtype=np.uint8)
        assert b.shape == (self.Hp, self.Wp)
        lib().orc_set_board(self.h, _p(b))

    def scalars(self):
        s = np.empty(6 + self.Q, np.int32)
        lib().orc_get_scalars(self.h, _p(s))
        return dict(x=int(s[0]), y=int(s[1]), active=int(s[2]), holder=int(s[3]),
                    has_swapped=bool(s[4]), game_over=bool(s[5]), queue=[int(v) for v in s[6:]])

    def active_matrix(self):
        n = C.c_int32()
        m = np.zeros(16, np.uint8)
        lib().orc_get_active_matrix(self.h, C.byref(n), _p(m))
        return m[: n.value * n.value].reshape(n.value, n.value).copy()

    def held_matrix(self):
        n = C.c_int32()
        m = np.zeros(16, np.uint8)
        lib().orc_get_held_matrix(self.h, C.byref(n), _p(m))
        return None if n.value == 0 else m[: n.value * n.value].reshape(n.value, n.value).copy()

    def set_active(self, idx, rot=0, x=None, y=None):
        s = self.scalars()
        lib().orc_set_active(self.h, int(idx), int(rot), s["x"] if x is None else int(x), s["y"] if y is None else int(y))

    def set_flags(self, has_swapped, game_over):
        lib().orc_set_flags(self.h, int(has_swapped), int(game_over))

    def set_holder(self, idx, rot=0):
        lib().orc_set_holder(self.h, -1 if idx is None else int(idx), int(rot))

    def set_queue(self, q):
        q = np.ascontiguousarray(q, dtype=np.int32)
        assert len(q) == self.Q
        lib().orc_set_queue(self.h, _p(q))


class OracleVec:
    """n independent oracle envs stepped with gymnasium's NEXT_STEP autoreset (bench baseline).
    bulk=True builds the envs inside the C library (no per-env Python object: `envs` is empty) -- a million envs in about a
    second instead of minutes; seed them with `seed_all` and reset with `reset_all`."""

    def __init__(self, n, bulk=False, **kw):
        self.n = n
        self._bulk = bulk
        if bulk:
            e = OracleEnv(**kw)
            self._proto = e
            self.envs = []
            self.ptrs = (C.c_void_p * n)()
            lib().orc_vec_create(C.byref(e.cfg), C.c_int64(n), self.ptrs)
        else:
            self.envs = [OracleEnv(**kw) for _ in range(n)]
            self.ptrs = (C.c_void_p * n)(*[e.h for e in self.envs])
            e = self.envs[0]
        self._init_buffers(n, e)

    def __del__(self):
        try:
            if self._bulk:
                lib().orc_vec_destroy(self.ptrs, C.c_int64(self.n))
        except Exception:
            pass

    def seed_all(self, seeds):
        """numpy-exact 7-bag streams: env i gets PCG64(SeedSequence(seeds[i])) (vectorised seeding, oracle/np_seed.py)."""
        from .np_seed import seed_words
        w = np.ascontiguousarray(seed_words(np.asarray(seeds, dtype=np.uint64)))
        lib().orc_vec_seed_words(self.ptrs, C.c_int64(self.n), _p(w))

    def reset_all(self, nthreads=0):
        lib().orc_vec_reset(self.ptrs, C.c_int64(self.n), int(nthreads))
        self.autoreset[:] = 0

    def _init_buffers(self, n, e):
        self.autoreset = np.zeros(n, np.uint8)
        self.board = np.empty((n, e.Hp, e.Wp), np.uint8)
        self.mask = np.empty((n, e.Hp, e.Wp), np.uint8)
        self.holder = np.empty((n, 4, 4 * e.holder_size), np.uint8)
        self.queue = np.empty((n, 4, 4 * e.Q), np.uint8)
        self.reward = np.empty(n, np.float32)
        self.terminated = np.empty(n, np.uint8)
        self.lines = np.empty(n, np.int32)
        self.feats = np.empty((n, 4 * e.W, e.W + 3), np.uint8)
        self.legal = np.ones((n, 4 * e.W), np.uint8)

    def step(self, actions, nthreads=0, autoreset=True):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        lib().orc_vec_step(self.ptrs, C.c_int64(self.n), _p(actions), _p(self.autoreset) if autoreset else None,
                           _p(self.board), _p(self.mask), _p(self.holder), _p(self.queue),
                           _p(self.reward), _p(self.terminated), _p(self.lines), int(nthreads))

    def grouped_step(self, actions, nthreads=0, autoreset=True):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        lib().orc_vec_grouped_step(self.ptrs, C.c_int64(self.n), _p(actions), _p(self.autoreset) if autoreset else None,
                                   _p(self.feats), _p(self.legal), _p(self.reward), _p(self.terminated),
                                   _p(self.lines), int(nthreads))

"""The reference's own base-env test suite (/root/reference/tests/test_base_env/**), restated against a small adapter so that
every case runs DIRECTLY on the CUDA env (tests/test_gpu_base_kats.py: Tetris(num_envs=1, randomizer_mode="numpy") -- the
numpy-exact 7-bag, so `reset(seed=42)` starts from the same piece, queue and position as the reference) and on the C oracle
(tests/test_oracle_base_kats.py).  Each case cites the reference test it follows (file:line) and asserts what that test
asserts; the reference pokes `env.unwrapped.board / x / y / active_tetromino`, the adapter maps those pokes onto
tg_set_state / the oracle's setters.
"""
import numpy as np

BASE = [np.array(m, np.uint8) for m in (
    [[0, 0, 0, 0], [1, 1, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]], [[1, 1], [1, 1]], [[0, 1, 0], [1, 1, 1], [0, 0, 0]],
    [[0, 1, 1], [1, 1, 0], [0, 0, 0]], [[1, 1, 0], [0, 1, 1], [0, 0, 0]], [[1, 0, 0], [1, 1, 1], [0, 0, 0]],
    [[0, 0, 1], [1, 1, 1], [0, 0, 0]])]
LEFT, RIGHT, DOWN, CW, CCW, HARD, SWAP, NOOP = range(8)      # ActionsMapping (mappings/actions.py:12-19)
ALIFE, GAME_OVER = 1.0, 0.0                                  # RewardsMapping (mappings/rewards.py:12-15)


class Adapter:
    """env.unwrapped of the reference, as far as its tests use it."""
    width, height, padding = 10, 20, 4
    width_padded, height_padded = 18, 24

    # x, y, board (locked cells incl. bedrock), game_over, has_swapped: properties with setters where the tests poke them
    def reset(self, seed=42): raise NotImplementedError
    def step(self, a): raise NotImplementedError          # -> obs dict, reward, terminated, truncated, {"lines_cleared"}
    def active_matrix(self): raise NotImplementedError    # n x n, id-valued
    def active_id(self): raise NotImplementedError
    def set_active(self, piece, rot=0): raise NotImplementedError
    def holder_ids(self): raise NotImplementedError
    def snapshot(self): raise NotImplementedError
    def restore(self, snap): raise NotImplementedError
    def fingerprint(self): raise NotImplementedError      # everything get_state() carries, as comparable values

    def fill(self, r0, r1, c0, c1, value=2):
        """env.unwrapped.board[r0:r1, c0:c1] = value (negative indices like numpy)"""
        b = self.board
        b[r0:r1, c0:c1] = value
        self.board = b

    def spawn_x(self):
        return self.width_padded // 2 - self.active_matrix().shape[0] // 2

    def score(self, n):                                    # Tetris.score (envs/tetris.py:621-630)
        return n * n * self.width


# ---- test_base_env_general.py -------------------------------------------------------------------------------------------------
def kat_observation_keys_and_shapes(mk):
    """test_base_env_general.py:13-35; test_base_env_action_space.py:1-6"""
    e = mk()
    obs = e.reset(seed=42)
    assert set(obs.keys()) == {"board", "active_tetromino_mask", "holder", "queue"}
    assert obs["board"].shape == (e.height_padded, e.width_padded)
    assert obs["active_tetromino_mask"].shape == (e.height_padded, e.width_padded)
    assert obs["holder"].shape == (4, 4) and obs["queue"].shape == (4, 16)
    assert e.action_space_n == 8


def kat_gravity_disabled(mk):
    """test_base_env_general.py:38-50"""
    e = mk(gravity=False)
    y0 = e.y
    e.step(NOOP)
    assert e.y == y0


def kat_clone_restore_consistency(mk):
    """test_base_env_general.py:141-176 (100 random actions; get_state / set_state round trip)"""
    e = mk()
    rng = np.random.default_rng(0)
    for t in range(100):
        if t == 0 or e.game_over:
            e.reset(seed=7 + t)
        snap = e.snapshot()
        a = int(rng.integers(0, 8))
        o1, r1, d1, _, i1 = e.step(a)
        post_a = e.fingerprint()
        e.restore(snap)
        o2, r2, d2, _, i2 = e.step(a)
        post_b = e.fingerprint()
        for k in o1:
            assert np.array_equal(o1[k], o2[k]), (t, k)
        assert r1 == r2 and d1 == d2 and i1 == i2
        assert all(np.array_equal(x, y) for x, y in zip(post_a, post_b)), t


# ---- test_base_env_reset.py ---------------------------------------------------------------------------------------------------
def kat_reset(mk):
    """test_base_env_reset.py:6-70"""
    e = mk()
    e.step(HARD); e.step(HARD)
    e.reset(seed=42)
    assert np.all(e.board[:e.height, e.padding:-e.padding] == 0)           # clean board after gameplay
    o1, o2 = e.reset(seed=42), e.reset(seed=42)
    for k in ("board", "queue", "holder"):
        assert np.array_equal(o1[k], o2[k])                                # same seed: deterministic
    e.fill(0, e.height, e.padding, -e.padding)
    e.step(HARD)
    assert e.game_over
    obs = e.reset(seed=42)
    assert not e.game_over and obs is not None
    obs, r, term, trunc, info = e.step(NOOP)
    assert obs is not None
    e.reset(seed=42)
    e.step(SWAP)
    assert len(e.holder_ids()) == 1
    e.reset(seed=42)
    assert len(e.holder_ids()) == 0                                        # reset clears the holder


# ---- actions/test_base_env_movement.py -------------------------------------------------------------------------------------------
def kat_movement(mk):
    """actions/test_base_env_movement.py:5-124"""
    e = mk()
    e.x = 5; e.step(RIGHT); assert e.x == 6
    e = mk(); e.x = 5; e.step(LEFT); assert e.x == 4
    e = mk(); e.y = 5; e.step(DOWN); assert e.y == 7                        # movement down (1) + gravity (1)
    e = mk(); e.x = e.width + e.padding - 1; e.step(RIGHT); assert e.x == e.width + e.padding - 1
    e = mk(); e.x = e.padding; e.step(LEFT); assert e.x == e.padding
    e = mk(); e.y = e.height - 1; e.step(DOWN); assert e.y == e.height - 1
    e = mk(); e.x = 5; e.fill(0, e.height, e.x + 1, e.x + 5); e.step(RIGHT); assert e.x == 5
    e = mk(); e.x = 5; e.fill(0, e.height, e.x - 4, e.x); e.step(LEFT); assert e.x == 5
    e = mk(); e.y = 5; e.fill(e.y + 1, e.y + 5, 0, e.width); e.step(DOWN); assert e.y == 5
    e = mk(gravity=False); e.y = 5; e.step(DOWN); assert e.y == 6          # only movement, no gravity
    e = mk(gravity=False)
    x0 = e.width_padded // 2
    e.x = x0
    for _ in range(3):
        e.step(LEFT)
    assert e.x == x0 - 3
    e = mk(); e.y = 0; e.step(NOOP); assert e.y == 1                        # gravity


# ---- actions/test_base_env_rotation.py -------------------------------------------------------------------------------------------
def kat_rotation(mk):
    """actions/test_base_env_rotation.py:9-92"""
    e = mk()
    m = e.active_matrix()
    e.step(CW)
    assert np.array_equal(np.rot90(m), e.active_matrix())
    e = mk()
    m = e.active_matrix()
    e.step(CCW)
    assert np.array_equal(np.rot90(m, -1), e.active_matrix())
    for a in (CW, CCW):                                                     # blocked by other tetrominoes
        e = mk()
        e.fill(0, e.height, e.x, e.x + 4)
        m = e.active_matrix()
        e.step(a)
        assert np.array_equal(m, e.active_matrix())
    e = mk(gravity=False)                                                   # blocked by the wall
    e.x = 0
    m = e.active_matrix()
    e.step(CW)
    assert np.array_equal(m, e.active_matrix())
    e = mk(gravity=False)
    m = e.active_matrix()
    for _ in range(4):
        e.step(CW)
    assert np.array_equal(m, e.active_matrix())                             # full 360
    e = mk(gravity=False)
    m = e.active_matrix()
    e.step(CW); e.step(CCW)
    assert np.array_equal(m, e.active_matrix())


# ---- actions/test_base_env_swap.py -----------------------------------------------------------------------------------------------
def kat_swap(mk):
    """actions/test_base_env_swap.py:4-99"""
    e = mk()
    first = e.active_id()
    assert len(e.holder_ids()) == 0
    e.step(SWAP)
    assert e.holder_ids() == [first] and e.active_id() is not None
    e = mk()
    first = e.active_id()
    e.step(SWAP); e.step(HARD); e.step(SWAP)
    assert e.active_id() == first                                           # swapping twice gets the piece back
    e = mk()
    e.step(SWAP)
    assert e.has_swapped is True
    after = e.active_id()
    e.step(SWAP)
    assert e.active_id() == after                                           # double swap blocked
    e = mk()
    e.step(SWAP); assert e.has_swapped is True
    e.step(HARD); assert e.has_swapped is False
    e = mk(gravity=False)
    e.y = 5; e.x = e.padding + 2
    e.step(HARD); e.step(SWAP); e.step(HARD)
    e.y = 5; e.x = e.padding + 3
    e.step(SWAP)
    assert e.x == e.spawn_x() and e.y == 0                                  # swap resets the position


# ---- actions/test_base_env_hard_drop.py, test_base_env_no_op.py ------------------------------------------------------------------
def kat_hard_drop_and_no_op(mk):
    """actions/test_base_env_hard_drop.py:6-73; actions/test_base_env_no_op.py:4-34"""
    e = mk()
    e.y = 0
    e.step(HARD)
    assert np.any(e.board[:e.height, e.padding:-e.padding] >= 2)
    e = mk()
    bottom = e.height - 4
    e.fill(bottom, e.height, e.padding, e.padding + 5)
    e.x = e.padding + 2; e.y = 0
    ph = e.active_matrix().shape[0]
    e.step(HARD)
    assert np.any(e.board[bottom - ph:bottom, e.padding:e.padding + 5] >= 2)   # lands on top of the existing blocks
    assert e.active_id() is not None
    e = mk()
    e.y = 5; e.x = e.padding + 2
    e.step(HARD)
    assert e.x == e.spawn_x() and e.y == 0
    e = mk(); e.x = 7; e.step(NOOP); assert e.x == 7
    e = mk(); e.y = 0; e.step(NOOP); assert e.y == 1
    e = mk(gravity=False)
    x0, y0 = e.x, e.y
    e.step(NOOP)
    assert (e.x, e.y) == (x0, y0)


# ---- reward/*.py -----------------------------------------------------------------------------------------------------------------
def kat_line_clear_and_scoring(mk):
    """reward/test_base_env_line_clear.py:11-89; reward/test_base_env_scoring.py:7-47"""
    e = mk()
    cleared = e.board.copy()
    e.fill(e.height - 4, -e.padding, e.padding, -e.padding - 1)             # everything but the last column, four rows
    assert np.any(e.board != cleared)
    e.set_active(0, 1)                                                      # vertical I = np.rot90(I)
    e.x = e.width + e.padding - 2
    obs, r, term, trunc, info = e.step(HARD)
    assert np.array_equal(e.board, cleared) and r == ((1 * 4) ** 2) * 10 + 1 and not term and not trunc
    assert e.x == e.spawn_x() and e.y == 0
    e = mk()
    e.fill(e.height - 2, e.height, e.padding, -e.padding - 2)
    e.set_active(1, 0)
    e.x = e.width + e.padding - 2
    obs, r, term, trunc, info = e.step(HARD)
    assert info["lines_cleared"] == 2 and r == (2 ** 2) * e.width + ALIFE
    e = mk()
    obs, r, term, trunc, info = e.step(HARD)
    assert info["lines_cleared"] == 0 and r == ALIFE and not term
    assert [e.score(n) for n in (1, 2, 3, 4)] == [10, 40, 90, 160]
    e = mk()
    e.fill(0, e.height, e.padding, -e.padding)
    obs, r, term, trunc, info = e.step(HARD)
    assert term and r == GAME_OVER


# ---- termination/test_base_env_game_over.py ----------------------------------------------------------------------------------------
def kat_game_over(mk):
    """termination/test_base_env_game_over.py:6-46"""
    for top in (2, 0):        # stack too high after the drop / piece already inside other tetrominoes
        e = mk()
        e.fill(top, e.height, e.padding, -(e.padding + 1))
        e.set_active(1, 0)
        e.x = e.width_padded // 2 - 1
        e.y = 0
        obs, r, term, trunc, info = e.step(HARD)
        assert term and r == GAME_OVER


ALL = [kat_observation_keys_and_shapes, kat_gravity_disabled, kat_clone_restore_consistency, kat_reset, kat_movement, kat_rotation,
       kat_swap, kat_hard_drop_and_no_op, kat_line_clear_and_scoring, kat_game_over]

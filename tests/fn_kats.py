"""Known-answer tests of the reference's FUNCTIONAL env, restated so that they run against BOTH the numpy oracle
(oracle/tetris_fn_oracle.py, CPU) and the CUDA facade (tg_fn_step through tetris_gymnasium_b200.envs.tetris_fn, GPU).

Every case follows one test function of /root/reference/tests/test_functional/ (file:line in its docstring) and asserts
what that test asserts.  The reference tests start from `reset(PRNGKey(42))`, i.e. from whatever piece the JAX key yields;
the assertions hold for any piece, so each case here runs for ALL 7 starting pieces (bags injected through `queue_fn`, the
reference's own hook).  The reference's core functions (collision, graviy_step, hard_drop, project_tetromino,
lock_active_tetromino, clear_filled_rows, check_game_over) are not separate entry points of the facade: their cases are
driven through `step` on constructed States (board / x / y / piece / rotation set with State.replace), which is how they
are reached in the reference's own step (envs/tetris_fn.py:161-273).

Each case returns a trace (list of plain values / arrays); tests/test_gpu_fn_kats.py additionally requires the GPU trace
to equal the oracle's trace element for element.
"""
import numpy as np

W, H, P, Q = 10, 20, 4, 7
SPAWN_X = (W + 2 * P) // 2 - 2      # core.get_initial_x_y: 4x4 matrices, board_width // 2 - 2 (functional/core.py:66-83)
BASE = [np.array(m, np.int8) for m in (
    [[0, 0, 0, 0], [1, 1, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]], [[1, 1], [1, 1]], [[0, 1, 0], [1, 1, 1], [0, 0, 0]],
    [[0, 1, 1], [1, 1, 0], [0, 0, 0]], [[1, 1, 0], [0, 1, 1], [0, 0, 0]], [[1, 0, 0], [1, 1, 1], [0, 0, 0]],
    [[0, 0, 1], [1, 1, 1], [0, 0, 0]])]


def matrix(piece, rot=0):
    """functional/tetrominoes.py:82-147: rot90(k = rot) of the base matrix, zero-padded to 4x4."""
    m = np.zeros((4, 4), np.int8)
    r = np.rot90(BASE[piece], k=rot)
    m[:r.shape[0], :r.shape[1]] = r
    return m


def empty_board():
    return np.pad(np.zeros((H, W), np.int8), ((0, P), (P, P)), constant_values=1)


def project(board, piece, rot, x, y):
    """core.project_tetromino (functional/core.py:103-121) on a numpy board: board[y:y+4, x:x+4] += matrix * id."""
    b = board.copy()
    b[y:y + 4, x:x + 4] += matrix(piece, rot) * np.int8(piece + 2)
    return b


def bag_for(first):
    return np.array([(first + i) % 7 for i in range(7)] * 4, np.uint8)


class Driver:
    """What a case may do with an env.  Implemented by OracleDriver (below) and GpuDriver (test_gpu_fn_kats.py)."""

    def get(self):      # -> dict(x, y, rotation, active, game_over, score, board int8[Hp, Wp], queue_index)
        raise NotImplementedError

    def set(self, **kw):
        raise NotImplementedError

    def step(self, a):  # -> (obs int8[H, W], reward float, terminated bool, lines int)
        raise NotImplementedError

    def obs(self):
        raise NotImplementedError


class OracleDriver(Driver):
    def __init__(self, first, gravity):
        from oracle.tetris_fn_oracle import FnOracle
        self.o = FnOracle(W, H, Q, gravity, seq=bag_for(first))
        self.reset_obs = self.o.reset()

    def get(self):
        o = self.o
        return dict(x=int(o.x), y=int(o.y), rotation=int(o.rot), active=int(o.active), game_over=bool(o.game_over), score=float(o.score),
                    board=o.board.copy(), queue_index=int(o.qidx))

    def set(self, **kw):
        names = {"rotation": "rot", "active": "active", "x": "x", "y": "y", "game_over": "game_over", "board": "board"}
        for k, v in kw.items():
            setattr(self.o, names[k], np.array(v, np.int8) if k == "board" else v)

    def step(self, a):
        obs, r, term, lines = self.o.step(a)
        return obs.copy(), float(r), bool(term), int(lines)

    def obs(self):
        return self.o.obs().copy()


# ---- cases -------------------------------------------------------------------------------------------------------------------
def kat_board_and_spawn(mk):
    """test_core/test_board.py:9-44 (TestCreateBoard, TestGetInitialXY) + test_env/test_reset.py:15-32"""
    tr = []
    for piece in range(7):
        d = mk(piece, True)
        s = d.get()
        assert s["board"].shape == (H + P, W + 2 * P)
        assert np.all(s["board"][:H, P:P + W] == 0)
        assert np.all(s["board"][:, :P] == 1) and np.all(s["board"][:, P + W:] == 1) and np.all(s["board"][H:, :] == 1)
        assert (s["x"], s["y"]) == (SPAWN_X, 0) and s["rotation"] == 0 and not s["game_over"] and s["score"] == 0.0
        assert s["active"] == piece
        tr += [s["board"], s["x"], s["y"]]
    return tr


def kat_collision(mk):
    """test_core/test_collision.py:8-55: no collision at the spawn position; the left bedrock, the bottom bedrock and a placed
    piece block a move (collision(x - 1) / collision(y + 1) inside update_state, envs/tetris_fn.py:186-205)."""
    tr = []
    d = mk(0, False)                       # I piece, row 1 of its matrix holds the cells
    s0 = d.get()
    d.step(5)                              # no_op without gravity: nothing collides at the spawn position, nothing moves
    assert d.get()["x"] == s0["x"] and d.get()["y"] == 0 and not d.get()["game_over"]
    d.set(x=P)                             # leftmost legal x: one step further left overlaps the bedrock
    d.step(0)
    assert d.get()["x"] == P
    d.set(x=P + 1)
    d.step(0)
    assert d.get()["x"] == P
    d.set(y=H - 2)                         # cells in board row H - 1: the bottom bedrock blocks a soft drop
    _, r, _, _ = d.step(2)
    assert d.get()["y"] == H - 2 and r == 0.0
    tr += [d.get()["x"], d.get()["y"]]
    # a placed T at (x, 10) blocks the same T coming down: free at y = 7 -> 8, blocked at y = 8 (test_gravity.py:29-43 numbers)
    d = mk(2, False)
    d.set(board=project(empty_board(), 2, 0, SPAWN_X, 10), y=7)
    d.step(2)
    assert d.get()["y"] == 8
    d.step(2)
    assert d.get()["y"] == 8
    # test_no_collision_adjacent: an O piece 4 rows above a placed O piece is free to move
    d = mk(1, False)
    d.set(board=project(empty_board(), 1, 0, SPAWN_X, 10), y=6)
    d.step(0)
    assert d.get()["x"] == SPAWN_X - 1
    tr += [d.get()["x"], d.get()["y"], d.get()["board"]]
    return tr


def kat_gravity(mk):
    """test_core/test_gravity.py:12-43 (graviy_step through step with gravity enabled, action no_op)"""
    tr = []
    for piece in range(7):
        d = mk(piece, True)
        obs, r, term, lines = d.step(5)
        assert d.get()["y"] == 1 and r == 0.0 and not term      # moves down on the empty board
        tr += [obs, d.get()["y"]]
    d = mk(0, True)
    d.set(y=H - 2)                        # blocked at the bottom: with gravity on, a blocked piece locks (should_lock)
    obs, r, term, lines = d.step(5)
    s = d.get()
    assert np.all(s["board"][H - 1, SPAWN_X:SPAWN_X + 4] == 2) and s["y"] == 0 and s["active"] == 1
    tr += [obs, s["board"], r]
    d = mk(2, True)
    d.set(board=project(empty_board(), 2, 0, SPAWN_X, 10), y=7)
    d.step(5)
    assert d.get()["y"] == 8              # free
    obs, r, term, lines = d.step(5)       # blocked by the placed piece -> locks at y = 8
    s = d.get()
    assert s["y"] == 0 and s["board"][9, SPAWN_X] == 4 and s["board"][8, SPAWN_X + 1] == 4
    tr += [obs, s["board"]]
    return tr


def kat_hard_drop(mk):
    """test_core/test_hard_drop.py:12-39 + test_env/test_hard_drop_action.py:10-44"""
    tr = []
    for piece in range(7):
        d = mk(piece, False)
        obs, r, term, lines = d.step(6)
        s = d.get()
        m = matrix(piece)
        lowest = max(i for i in range(4) if m[i].any())
        land_y = H - 1 - lowest
        assert r == 2 * land_y and r > 0                        # reward is twice the distance; positive
        assert np.any(s["board"][:H, P:P + W] > 0) and s["y"] == 0      # piece locked, new piece spawned at the top
        assert np.array_equal(s["board"][land_y:land_y + 4, SPAWN_X:SPAWN_X + 4][m > 0], np.full(4, piece + 2, np.int8))
        assert land_y >= H - 4                                   # "should be near the bottom"
        cells = int((s["board"][:H, P:P + W] > 0).sum())
        obs, r2, term, lines = d.step(6)                         # lands on the existing piece
        assert int((d.get()["board"][:H, P:P + W] > 0).sum()) == cells + 4
        tr += [s["board"], r, d.get()["board"], r2]
    d = mk(2, False)
    d.set(board=project(empty_board(), 2, 0, SPAWN_X, 15))
    d.step(6)
    rows = np.flatnonzero((d.get()["board"][:H, P:P + W] > 0).any(1))
    assert rows.min() < 15                                       # stops on the existing piece
    tr += [d.get()["board"]]
    return tr


def kat_line_clear(mk):
    """test_core/test_line_clear.py:10-71 + test_core/test_lock.py:15-64 (lock_active_tetromino -> clear_filled_rows -> score),
    driven through a hard drop of an O piece far from the prepared rows' gap-free cells."""
    tr = []
    for k, want in ((0, 0), (1, 100), (2, 300), (4, 800)):
        d = mk(1, False)                    # O piece
        b = empty_board()
        for i in range(k):
            b[H - 1 - i, P:P + W] = 2
        if k:
            b[H - 1 - k, P] = 5             # marker above the full rows: shifts down by k
        d.set(board=b)
        obs, r, term, lines = d.step(6)
        s = d.get()
        drop = r - want
        assert lines == k and drop == 2 * (H - 2 - k)      # reward = drop distance * 2 + score(lines) (test_lock.py:48-64)
        if k:
            assert s["board"][H - 1, P] == 5 and np.all(s["board"][H - 1, P + 1:SPAWN_X] == 0)
        assert np.all(s["board"][H - 2:H, SPAWN_X:SPAWN_X + 2] == 3)    # the O piece came down with the stack
        tr += [s["board"], r, lines]
    d = mk(1, False)                        # partial row is not cleared
    b = empty_board()
    b[H - 1, P:P + W - 1] = 2
    d.set(board=b)
    _, r, _, lines = d.step(6)
    assert lines == 0 and np.all(d.get()["board"][H - 1, P:P + W - 1] == 2)
    tr += [d.get()["board"], r]
    return tr


def kat_line_clear_row0_occupied(mk):
    """The `jnp.take(sub_board, indices, axis=0, fill_value=0)` question (functional/core.py:203-214): cleared rows get index
    -config.height.  jnp.take's default mode "fill" first wraps negative indices numpy-style (index + axis_size), so -H is
    row 0 -- a VALID index, not an out-of-bounds one: the n new top rows are copies of the old row 0, not zeros.  Identical
    whenever row 0 is empty; with a cell in row 0 the copies carry it.  Chosen answer: the published jnp.take semantics
    (copies of row 0); jax is not installed here, so this case is pinned to that reading, not to an execution."""
    d = mk(1, False)
    b = empty_board()
    b[H - 1, P:P + W] = 2
    b[H - 2, P:P + W] = 2
    b[0, P + W - 1] = 6                      # a cell in row 0, away from the spawn columns
    d.set(board=b)
    obs, r, term, lines = d.step(6)
    s = d.get()
    assert lines == 2
    assert s["board"][0, P + W - 1] == 6 and s["board"][1, P + W - 1] == 6     # two new top rows = copies of old row 0
    assert s["board"][2, P + W - 1] == 6                                      # old row 0 itself moved down by two
    assert int((s["board"][:H, P + W - 1] > 0).sum()) == 3
    return [s["board"], r, lines, obs]


def kat_projection(mk):
    """test_core/test_projection.py:12-54: locking writes id * matrix where the matrix is set, keeps what was there, and leaves
    the matrix' zero cells alone (observed on the board after a hard drop)."""
    tr = []
    d = mk(0, False)
    d.step(6)                               # I piece to the floor (row H - 1)
    b1 = d.get()["board"]
    assert np.array_equal(b1[H - 1, SPAWN_X:SPAWN_X + 4], np.full(4, 2, np.int8))
    d.step(6)                               # next piece of the injected bag (O) on top
    b2 = d.get()["board"]
    assert np.array_equal(b2[H - 1, SPAWN_X:SPAWN_X + 4], np.full(4, 2, np.int8))          # preserved
    assert np.all(b2[H - 3:H - 1, SPAWN_X:SPAWN_X + 2] == 3)
    assert np.all(b2[:H - 3, P:P + W] == 0) and np.all(b2[H - 3:H - 1, SPAWN_X + 2:P + W] == 0)   # zero cells untouched
    tr += [b1, b2]
    return tr


def kat_game_over(mk):
    """test_core/test_game_over.py:13-27 (check_game_over at spawn) + test_env/test_step.py:16-25 + test_observations.py:44-60"""
    tr = []
    for piece in range(7):
        d = mk(piece, False)
        nxt = (piece + 1) % 7
        # the NEXT piece's spawn cells are filled; the falling piece starts below them, in the left corner
        d.set(board=project(empty_board(), nxt, 0, SPAWN_X, 0), x=P, y=5)
        obs, r, term, lines = d.step(6)
        assert term and d.get()["game_over"]
        assert not np.any(obs == -1)                           # no active piece in the observation once the game is over
        frozen = d.get()
        obs2, r2, term2, lines2 = d.step(0)                    # a finished game is frozen
        s = d.get()
        assert term2 and r2 == 0.0 and np.array_equal(s["board"], frozen["board"]) and s["x"] == frozen["x"]
        tr += [obs, frozen["board"], r, obs2]
    d = mk(3, True)                                            # forcing game_over on a live state (test_step.py:16-25)
    d.set(game_over=True)
    before = d.get()["board"]
    obs, r, term, lines = d.step(0)
    assert term and r == 0.0 and np.array_equal(d.get()["board"], before) and not np.any(obs == -1)
    tr += [obs]
    return tr


def kat_movement(mk):
    """test_env/test_movement.py:9-73"""
    tr = []
    for piece in range(7):
        d = mk(piece, False)
        d.step(0)
        assert d.get()["x"] == SPAWN_X - 1
        d.step(1); d.step(1)
        assert d.get()["x"] == SPAWN_X + 1
        _, r, _, _ = d.step(2)
        assert d.get()["y"] == 1 and r == 1.0                  # soft drop: one point per cell
        for _ in range(3):
            d.step(1)
        assert d.get()["x"] == min(SPAWN_X + 4, d.get()["x"])  # three consecutive moves (clamped by the wall for wide pieces)
        for _ in range(20):
            d.step(0)
        x_left = d.get()["x"]
        d.step(0)
        assert d.get()["x"] == x_left                          # blocked at the left border
        for _ in range(20):
            d.step(1)
        x_right = d.get()["x"]
        d.step(1)
        assert d.get()["x"] == x_right and x_right > x_left    # blocked at the right border
        for _ in range(30):
            d.step(2)
        y_bot = d.get()["y"]
        _, r, _, _ = d.step(2)
        assert d.get()["y"] == y_bot and r == 0.0              # soft drop blocked at the bottom
        tr += [x_left, x_right, y_bot]
        d = mk(piece, False)                                   # blocked by a placed piece (test_movement.py:58-73)
        d.step(6)
        prev = -1
        for _ in range(30):
            prev = d.get()["y"]
            _, _, term, _ = d.step(2)
            if term or d.get()["y"] == prev:
                break
        assert d.get()["y"] == prev or term
        tr += [d.get()["y"], d.get()["board"]]
    return tr


def kat_rotation(mk):
    """test_env/test_rotation.py:9-43 (3 = counter-clockwise, 4 = clockwise)"""
    tr = []
    for piece in range(7):
        d = mk(piece, False)
        d.set(y=2)                                             # away from the ceiling: every rotation is free
        d.step(4)
        assert d.get()["rotation"] == 1
        d.step(3)
        assert d.get()["rotation"] == 0
        d.step(3)
        assert d.get()["rotation"] == 3
        for _ in range(4):
            d.step(4)
        assert d.get()["rotation"] == 3                        # full 360 degree cycle
        obs = d.obs()
        assert (obs == -1).sum() == 4
        for _ in range(15):
            d.step(0)
        d.step(4)                                              # at the wall: may succeed or be blocked, stays a valid rotation
        assert 0 <= d.get()["rotation"] <= 3
        tr += [obs, d.get()["rotation"], d.get()["x"], d.obs()]
    return tr


def kat_no_op(mk):
    """test_env/test_no_op.py:7-28"""
    tr = []
    for piece in range(7):
        d = mk(piece, False)
        s0 = d.get()
        obs, r, term, lines = d.step(5)
        s1 = d.get()
        assert (s1["x"], s1["y"], s1["rotation"]) == (s0["x"], s0["y"], s0["rotation"]) and r == 0.0
        g = mk(piece, True)
        g.step(5)
        assert g.get()["x"] == s0["x"] and g.get()["y"] >= s0["y"]
        tr += [obs, g.get()["y"]]
    return tr


def kat_observations(mk):
    """test_env/test_observations.py:10-60 + test_env/test_step.py:27-35"""
    tr = []
    for piece in range(7):
        d = mk(piece, True)
        obs = d.obs()
        assert obs.shape == (H, W) and obs.dtype == np.int8 and set(np.unique(obs).tolist()) <= {-1, 0, 1} and (obs == -1).sum() == 4
        s0 = d.get()
        o1, r, term, lines = d.step(5)
        assert o1.shape == (H, W) and r == d.get()["score"] - s0["score"]       # reward equals the score difference
        d.step(6)
        o2 = d.obs()
        assert (o2 == 1).sum() == 4 and (o2 == -1).sum() == 4                  # locked cells read 1, the active piece -1
        tr += [obs, o1, o2]
    return tr


def kat_bag_queue(mk):
    """test_queue.py:23-71 with injected bags: the queue hands out its 7 entries in order, then refills (index 1 after the first
    piece of the new bag)."""
    d = mk(3, False)
    seen = [d.get()["active"]]
    for i in range(7):
        d.set(x=P + (i % 3) * 3)             # spread the drops so that the stack stays low
        d.step(6)
        seen.append(d.get()["active"])
    assert seen[:7] == [(3 + i) % 7 for i in range(7)] and set(seen[:7]) == set(range(7))
    assert seen[7] == 3 and d.get()["queue_index"] == 1                        # refill after exhaustion
    return [seen]


ALL = [kat_board_and_spawn, kat_collision, kat_gravity, kat_hard_drop, kat_line_clear, kat_line_clear_row0_occupied, kat_projection,
       kat_game_over, kat_movement, kat_rotation, kat_no_op, kat_observations, kat_bag_queue]

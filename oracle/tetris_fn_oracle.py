"""numpy restatement of the reference FUNCTIONAL env (TEST INFRASTRUCTURE ONLY).

Follows tetris_gymnasium/envs/tetris_fn.py:137-413, functional/core.py:46-302, functional/queue.py:20-67 and
functional/tetrominoes.py:42-165 line by line on plain numpy (small cases only).

PARITY UNPINNED for key-derived piece sequences: jax / chex are not installed in the build container, so the
reference functional env cannot be executed here, and jax.random.permutation (threefry) is not restated.
Bags are therefore INJECTED (bag k = seq[k*Q:(k+1)*Q]), which is the reference's own hook (`queue_fn`,
`create_queue_fn`, or overwriting state.queue).  What IS pinned are the reference tests' RNG-free known answers
(tests/test_oracle_fn.py): score table (tests/test_functional/test_core/test_scoring.py:11-17), line-clear
counts and shifting (test_core/test_line_clear.py:14-71), step-when-game-over no-op (test_env/test_step.py:16-25),
observation value set / shape (test_env/test_observations.py).
clear_filled_rows (SURVEY 3.4 open point, now decided): the reference gathers rows with jnp.take(..., fill_value=0) and
gives cleared rows the index -H.  jnp.take's default mode "fill" wraps negative indices numpy-style before the bounds
check (jax/_src/numpy/lax_numpy.py `_take`: `indices = where(indices < 0, indices + axis_size, indices)`), so -H is row 0,
in bounds: the n new top rows are copies of the OLD ROW 0, not zeros.  Identical whenever row 0 is empty.  This
restatement and the CUDA facade follow that published semantics (tests/fn_kats.py::kat_line_clear_row0_occupied); it is
pinned to the documented behaviour of jax 0.5.3, not to an execution (jax is not installed here).
"""
import numpy as np

P = 4
_BASE = [
    [[0, 0, 0, 0], [1, 1, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]], [[1, 1], [1, 1]], [[0, 1, 0], [1, 1, 1], [0, 0, 0]],
    [[0, 1, 1], [1, 1, 0], [0, 0, 0]], [[1, 1, 0], [0, 1, 1], [0, 0, 0]], [[1, 0, 0], [1, 1, 1], [0, 0, 0]],
    [[0, 0, 1], [1, 1, 1], [0, 0, 0]]]
MATRICES = np.zeros((7, 4, 4, 4), np.int8)   # functional/tetrominoes.py:82-147
for _p, _m in enumerate(_BASE):
    for _r in range(4):
        _rm = np.rot90(np.array(_m, np.int8), k=_r)
        MATRICES[_p, _r, : _rm.shape[0], : _rm.shape[1]] = _rm
IDS = np.arange(2, 9)


def create_board(W, H):  # core.create_board :46-63
    return np.pad(np.zeros((H, W), np.int8), ((0, P), (P, P)), constant_values=1)


def collision(board, m, x, y):  # core.collision :86-100
    return bool(np.any((board[y:y + 4, x:x + 4] > 0) & (m > 0)))


def score(rows):  # core.score :124-146
    return 800 if rows == 4 else (rows * 200 - 100 if rows > 0 else 0)


def clear_filled_rows(board, W, H):  # core.clear_filled_rows :185-227
    sub = board[:H, P:P + W]
    filled = np.all(sub > 0, axis=1)
    n = int(filled.sum())
    if n == 0:
        return board, 0
    # indices = sort(where(filled, -H, arange(H))); jnp.take(sub, indices, axis=0, fill_value=0): in the default mode "fill"
    # jnp.take first wraps negative indices numpy-style (index + axis_size), so -H addresses row 0 -- a valid index: the n
    # new top rows are COPIES OF THE OLD ROW 0, not zeros (identical whenever row 0 is empty)
    new_sub = np.concatenate([np.repeat(sub[0:1], n, axis=0), sub[~filled]], axis=0)
    return np.pad(new_sub, ((0, P), (P, P)), constant_values=1), n


class FnOracle:
    def __init__(self, W=10, H=20, Q=7, gravity=True, seq=None):
        self.W, self.H, self.Q, self.gravity = W, H, Q, gravity
        self.seq = np.asarray(seq)
        self.bagno = 0

    def _new_bag(self):
        L = len(self.seq)
        q = np.array([self.seq[(self.bagno * self.Q + i) % L] for i in range(self.Q)], np.int32)
        self.bagno += 1
        return q

    def reset(self):  # tetris_fn.reset :318-367
        self.board = create_board(self.W, self.H)
        self.bagno = 0
        self.queue = self._new_bag()
        self.active, self.qidx = int(self.queue[0]), 1
        self.rot, self.x, self.y = 0, (self.W + 2 * P) // 2 - 2, 0
        self.game_over, self.score = False, np.float32(0)
        return self.obs()

    def obs(self):  # get_observation :137-158
        b = (self.board > 0).astype(np.int8)
        if not self.game_over:
            b = b.copy()
            b[self.y:self.y + 4, self.x:self.x + 4] += MATRICES[self.active, self.rot] * np.int8(-1)
        return b[:self.H, P:P + self.W]

    def step(self, a):  # step :276-315 + update_state :161-273
        old = self.score
        lines = 0
        if not self.game_over:
            m = MATRICES[self.active, self.rot]
            drop = 0
            if a == 0 and not collision(self.board, m, self.x - 1, self.y):
                self.x -= 1
            elif a == 1 and not collision(self.board, m, self.x + 1, self.y):
                self.x += 1
            elif a == 2 and not collision(self.board, m, self.x, self.y + 1):
                self.y += 1
                drop = 1
            elif a in (3, 4):
                nr = (self.rot + (1 if a == 4 else -1)) % 4
                if not collision(self.board, MATRICES[self.active, nr], self.x, self.y):
                    self.rot, m = nr, MATRICES[self.active, nr]
            elif a == 6:  # core.hard_drop :230-251
                ny = self.y
                while not collision(self.board, m, self.x, ny + 1):
                    ny += 1
                drop = 2 * (ny - self.y)
                self.y = ny
            yg = self.y
            if self.gravity and not collision(self.board, m, self.x, self.y + 1):
                yg = self.y + 1
            should_lock = (yg == self.y) and self.gravity
            self.y = yg
            lock = 0
            if should_lock or a == 6:  # place_active_tetromino :370-413
                self.board = self.board.copy()
                self.board[self.y:self.y + 4, self.x:self.x + 4] += m * np.int8(IDS[self.active])
                self.board, lines = clear_filled_rows(self.board, self.W, self.H)
                lock = score(lines)
                if self.qidx >= self.Q:  # queue.bag_queue_get_next_element :38-67
                    self.queue = self._new_bag()
                    self.active, self.qidx = int(self.queue[0]), 1
                else:
                    self.active, self.qidx = int(self.queue[self.qidx]), self.qidx + 1
                self.rot, self.x, self.y = 0, (self.W + 2 * P) // 2 - 2, 0
                self.game_over = collision(self.board, MATRICES[self.active, 0], self.x, self.y)  # check_game_over
            self.score = np.float32(self.score + np.float32(drop + lock))
        return self.obs(), np.float32(self.score - old), self.game_over, lines

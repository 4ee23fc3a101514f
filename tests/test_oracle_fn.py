"""CPU: the numpy restatement of the functional env against the reference tests' RNG-free known answers
(jax is absent, so key-derived sequences stay unpinned -- see the oracle's header)."""
import numpy as np

from oracle.tetris_fn_oracle import MATRICES, FnOracle, clear_filled_rows, create_board, score


def test_score_table():  # tests/test_functional/test_core/test_scoring.py:11-17
    assert [score(k) for k in range(5)] == [0, 100, 300, 500, 800]


def test_line_clear_known_answers():  # tests/test_functional/test_core/test_line_clear.py:14-71
    b = create_board(10, 20)
    assert clear_filled_rows(b.copy(), 10, 20)[1] == 0
    for k in (1, 2, 4):
        c = b.copy()
        c[20 - k:20, 4:14] = 2
        nb, n = clear_filled_rows(c, 10, 20)
        assert n == k and np.all(nb[:20, 4:14] == 0)
    c = b.copy(); c[19, 4:13] = 2
    assert clear_filled_rows(c, 10, 20)[1] == 0
    c = b.copy(); c[19, 4:14] = 2; c[18, 4] = 3      # marker above a full row shifts down by one
    nb, n = clear_filled_rows(c, 10, 20)
    assert n == 1 and nb[19, 4] == 3 and nb[18, 4] == 0


def test_step_when_game_over_is_noop_and_obs_values():  # test_env/test_step.py:16-25, test_observations.py
    o = FnOracle(seq=np.arange(7))
    obs = o.reset()
    assert obs.shape == (20, 10) and obs.dtype == np.int8 and set(np.unique(obs)) <= {-1, 0, 1} and (obs == -1).sum() == 4
    o.game_over = True
    board = o.board.copy()
    obs, r, term, lines = o.step(0)
    assert term and r == 0.0 and np.array_equal(o.board, board) and (obs == -1).sum() == 0


def test_matrices_match_numpy_env_geometry():
    # same rot90(k=r) geometry as the NumPy env; I piece rows as bit masks
    rows = lambda m: [int(sum(int(v) << j for j, v in enumerate(r))) for r in m]
    assert rows(MATRICES[0, 0]) == [0, 15, 0, 0] and rows(MATRICES[0, 1]) == [2, 2, 2, 2]
    assert rows(MATRICES[6, 3])[:3] == [2, 2, 6]


def test_hard_drop_scores_two_per_cell_and_locks():
    o = FnOracle(gravity=False, seq=np.array([1, 1, 1, 1, 1, 1, 1]))
    o.reset()
    obs, r, term, lines = o.step(6)       # O piece from y=0 to the floor: 18 cells
    assert r == 36.0 and not term and lines == 0 and o.board[18:20, 7:9].tolist() == [[3, 3], [3, 3]]

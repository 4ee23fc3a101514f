"""Functional env facade -- mirrors tetris_gymnasium/envs/tetris_fn.py (reset :318-367, step :276-315,
batched_step :416-435, batched_reset :438-467, ACTION_ID_TO_NAME :470-478) on torch CUDA tensors.

Same signatures: `reset(tetrominoes, key, config, create_queue_fn, queue_fn) -> (key, state, obs)` and
`step(tetrominoes, state, action, config, queue_fn) -> (state, obs, reward, terminated, info)`; the
`batched_*` variants take a leading batch axis and keyword-only `config`.  The game rules of this path
differ from the NumPy env (7 actions, no holder, queue == bag, score-delta reward) and are
executed by `tg_fn_step` (csrc/tg_fn.cuh).

`queue_fn` / `create_queue_fn`: the reference takes JAX callables; here they select the bag source:
`None`/`"bag"` = device Philox permutations keyed by `state.rng_key` (NOT bit-compatible with
jax.random.permutation), `"uniform"` (or the reference's uniform-queue callables by name) = the uniform queue of
functional/queue.py:71-119, or a uint8 tensor `[B, L]` of injected bags (bag k = seq[:, k*Q:(k+1)*Q]).
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..functional.core import EnvConfig, State
from ..functional.tetrominoes import TETROMINOES, Tetrominoes  # noqa: F401

ACTION_ID_TO_NAME = {0: "move_left", 1: "move_right", 2: "move_down", 3: "rotate_counterclockwise",
                     4: "rotate_clockwise", 5: "do_nothing", 6: "hard_drop"}
_S = 9  # TG_FN_SCALARS


def _dev(x=None):
    if torch.is_tensor(x) and x.is_cuda:
        return x.device
    if not torch.cuda.is_available():
        raise RuntimeError("tetris_gymnasium_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _fields(st: State):
    return (st.rng_key, st.active_tetromino, st.rotation, st.x, st.y, st.queue, st.queue_index, st.game_over, st.score)


def _pack(state: State, Q: int) -> torch.Tensor:
    """State -> the i32[B, 9 + Q] record array tg_fn_step reads.  A State that came out of `_unpack` and whose field
    tensors are still the same objects at the same version (no attribute assignment, no in-place edit, no replace())
    carries its record array: the usual `state = step(state)` loop packs nothing."""
    tag = getattr(state, "_tg_packed", None)
    if tag is not None:
        sc, sig = tag
        if sc.shape[1] == _S + Q and all(id(t) == i and t._version == v for t, (i, v) in zip(_fields(state), sig)):
            return sc
    B = state.board.shape[0]
    sc = torch.empty((B, _S + Q), dtype=torch.int32, device=state.board.device)
    sc[:, 0] = state.active_tetromino
    sc[:, 1] = state.rotation
    sc[:, 2] = state.x
    sc[:, 3] = state.y
    sc[:, 4] = state.queue_index
    sc[:, 5] = state.game_over.to(torch.int32)
    sc[:, 6] = state.score.to(torch.float32).view(torch.int32)
    sc[:, 7:9] = state.rng_key.to(torch.int64).to(torch.int32)
    sc[:, _S:] = state.queue
    return sc


def _unpack(board: torch.Tensor, sc: torch.Tensor) -> State:
    st = State(rng_key=sc[:, 7:9].to(torch.int64) & 0xFFFFFFFF, board=board, active_tetromino=sc[:, 0], rotation=sc[:, 1],
               x=sc[:, 2], y=sc[:, 3], queue=sc[:, _S:], queue_index=sc[:, 4], game_over=sc[:, 5] != 0,
               score=sc[:, 6].contiguous().view(torch.float32))
    st._tg_packed = (sc, tuple((id(t), t._version) for t in _fields(st)))
    return st


UNIFORM = "uniform"   # queue selector: functional/queue.py:71-119 (create_uniform_queue / uniform_queue_get_next_element)


def _is_uniform(queue_fn):
    return (isinstance(queue_fn, str) and queue_fn == UNIFORM) or getattr(queue_fn, "__name__", "") in (
        "create_uniform_queue", "uniform_queue_get_next_element")


def _seq(queue_fn, dev):
    """None / "bag" / the bag-queue callables -> Philox permutations; "uniform" / the uniform-queue callables -> uniform
    queue; a uint8 array [B, L] -> injected bags."""
    if queue_fn is None or isinstance(queue_fn, str) or callable(queue_fn):
        return UNIFORM if _is_uniform(queue_fn) else None
    s = torch.as_tensor(np.asarray(queue_fn) if not torch.is_tensor(queue_fn) else queue_fn)
    return s.to(dev, torch.uint8).contiguous()


def _call(config: EnvConfig, board_in, sc_in, actions, seq):
    if config.padding != 4:
        raise ValueError("padding must be 4 (the tetromino matrices are 4x4)")
    L = _lib.load()
    dev = board_in.device
    B = board_in.shape[0]
    board_out = torch.empty_like(board_in)
    sc_out = torch.empty_like(sc_in)
    obs = torch.empty((B, config.height, config.width), dtype=torch.int8, device=dev)
    reward = torch.empty(B, dtype=torch.float32, device=dev)
    term = torch.empty(B, dtype=torch.uint8, device=dev)
    lines = torch.empty(B, dtype=torch.int32, device=dev)
    uniform = isinstance(seq, str)
    if uniform:
        seq = None
    if seq is not None:
        assert seq.shape[0] == B
    with torch.cuda.device(dev):
        rc = L.tg_fn_step(config.width, config.height, config.queue_size, int(bool(config.gravity_enabled)), B,
                          board_in.data_ptr(), sc_in.data_ptr(), actions.data_ptr() if actions is not None else None,
                          seq.data_ptr() if seq is not None else None, seq.shape[1] if seq is not None else (-1 if uniform else 0),
                          board_out.data_ptr(), sc_out.data_ptr(), obs.data_ptr(), reward.data_ptr(), term.data_ptr(),
                          lines.data_ptr(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc)
    return board_out, sc_out, obs, reward, term.view(torch.bool), lines


def batched_reset(tetrominoes, keys, *, config: EnvConfig, create_queue_fn=None, queue_fn=None, batch_size: int = 1):
    """keys: [B, 2] (uint32 values).  Returns (keys, states, observations) like the reference (:438-467)."""
    dev = _dev(keys)
    keys = torch.as_tensor(np.asarray(keys) if not torch.is_tensor(keys) else keys).to(dev).to(torch.int64).reshape(-1, 2)
    B = keys.shape[0]
    Hp, Wp = config.height + config.padding, config.width + 2 * config.padding
    board = torch.zeros((B, Hp, Wp), dtype=torch.int8, device=dev)
    sc = torch.zeros((B, _S + config.queue_size), dtype=torch.int32, device=dev)
    sc[:, 7] = keys[:, 0].to(torch.int32)
    sc[:, 8] = keys[:, 1].to(torch.int32)
    seq = _seq(create_queue_fn if create_queue_fn is not None else queue_fn, dev)
    board, sc, obs, _, _, _ = _call(config, board, sc, None, seq)
    # the reference returns split(key)[0] as the caller's new key; ours advances the second word
    new_keys = torch.stack([keys[:, 0], (keys[:, 1] + 1) & 0xFFFFFFFF], dim=1)
    return new_keys, _unpack(board, sc), obs


def batched_step(tetrominoes, states: State, actions, *, config: EnvConfig, queue_fn=None):
    """Vectorised step (:416-435): (states, observations, rewards, terminated, info)."""
    dev = states.board.device
    a = torch.as_tensor(np.asarray(actions) if not torch.is_tensor(actions) else actions).to(dev, torch.int32).reshape(-1).contiguous()
    board, sc, obs, reward, term, lines = _call(config, states.board.contiguous(), _pack(states, config.queue_size), a, _seq(queue_fn, dev))
    return _unpack(board, sc), obs, reward, term, {"lines_cleared": lines}


def reset(tetrominoes, key, config: EnvConfig, create_queue_fn=None, queue_fn=None):
    """Single-env reset (:318-367); tensors keep a leading axis of size 1."""
    key = torch.as_tensor(np.asarray(key) if not torch.is_tensor(key) else key).reshape(1, 2)
    keys, state, obs = batched_reset(tetrominoes, key, config=config, create_queue_fn=create_queue_fn, queue_fn=queue_fn)
    return keys[0], state, obs[0]


def step(tetrominoes, state: State, action, config: EnvConfig, queue_fn=None):
    """Single-env step (:276-315)."""
    state, obs, reward, term, info = batched_step(tetrominoes, state, [int(action)], config=config, queue_fn=queue_fn)
    return state, obs[0], reward[0], term[0], {"lines_cleared": info["lines_cleared"][0]}

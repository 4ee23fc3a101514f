"""Batched Tetris env on B200 -- host-side mirror of the reference `Tetris(gym.Env)`.

Reference: tetris_gymnasium/envs/tetris.py (constructor :77-91, step :203-272, reset :274-307,
_get_obs :566-615, get_state/set_state :681-708).  Same constructor options, same action /
reward mappings, same observation-dict layout -- with a leading env axis: every array is a torch
CUDA tensor `[num_envs, ...]`.  All game logic runs in libtetris_b200.so (hand-written sm_100a
CUDA behind a C ABI, include/tetris_b200.h); torch only owns memory and streams.  No CPU path.
"""
from dataclasses import fields
from typing import Any, Optional

import ctypes as C
import os

import numpy as np
import torch

from .. import _lib
from ..components.tetromino import Pixel, Tetromino
from ..mappings.actions import ActionsMapping
from ..mappings.rewards import RewardsMapping

PADDING = 4  # max tetromino matrix dim (reference envs/tetris.py:130)
_NO_OBS = _lib.TgObs(None, None, None, None)   # all-NULL tg_obs: the step skips the observation dict


class _Space:
    """Tiny stand-in used when gymnasium is not importable (shape / dtype / n only)."""

    def __init__(self, shape=(), dtype=np.uint8, low=0, high=0, n=None):
        self.shape, self.dtype, self.low, self.high, self.n = tuple(shape), np.dtype(dtype), low, high, n

    def contains(self, x):
        if self.n is not None:
            return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n
        return True


def _spaces(env):
    try:
        import gymnasium as gym
        from gymnasium.spaces import Box, Discrete

        mk_box = lambda hi, shape: Box(low=0, high=hi, shape=shape, dtype=np.uint8)  # noqa: E731
        disc = Discrete(8)
        dct = gym.spaces.Dict
    except Exception:  # gymnasium absent: attribute-compatible stand-ins
        mk_box = lambda hi, shape: _Space(shape, np.uint8, 0, hi)  # noqa: E731
        disc = _Space(n=8, dtype=np.int64)
        dct = dict
    n_pix = 2 + (7 if env.tetrominoes is None else len(env.tetrominoes))  # len(self.pixels) (reference envs/tetris.py:127)
    obs = dct({
        "board": mk_box(n_pix, (env.height_padded, env.width_padded)),
        "active_tetromino_mask": mk_box(1, (env.height_padded, env.width_padded)),
        "holder": mk_box(n_pix, (env.padding, env.padding * env.holder_size)),
        "queue": mk_box(n_pix, (env.padding, env.padding * env.queue_size)),
    })
    return obs, disc


class Tetris:
    """`num_envs` independent Tetris games stepped by one CUDA kernel launch.

    Positional / keyword options are the reference's (envs/tetris.py:77-91); the keyword-only
    ones are ours:

      num_envs        number of envs on this device
      device          torch device (default: current CUDA device)
      queue_size      visible queue length (reference TetrominoQueue default 4; EnvConfig.queue_size)
      padding         must be 4 (derived in the reference, EnvConfig field in the functional env)
      autoreset_mode  "next_step" (gymnasium 1.x vector default) | "same_step" | "disabled"
      randomizer_mode "philox" (device-native streams) | "numpy" (bit-exact numpy streams: PCG64 +
                      Generator.shuffle / Generator.integers) | "sequence" (injected piece streams, `piece_sequences`)
      randomizer      None / "bag" / BagRandomizer = 7-bag (reference default); "true" / TrueRandomizer = uniform draws
                      (components/tetromino_randomizer.py:105-136)
      env_id_offset   global id of env 0 (multi-GPU sharding keeps Philox streams independent of #GPUs)
    """

    metadata = {"render_modes": ["rgb_array", "ansi"], "render_fps": 1}

    # the reference's class attributes (envs/tetris.py:45-75): users build custom sets from copies of these
    BASE_PIXELS = [Pixel(0, [0, 0, 0]), Pixel(1, [128, 128, 128])]
    TETROMINOES = [
        Tetromino(0, [0, 240, 240], np.array([[0, 0, 0, 0], [1, 1, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]], dtype=np.uint8)),
        Tetromino(1, [240, 240, 0], np.array([[1, 1], [1, 1]], dtype=np.uint8)),
        Tetromino(2, [160, 0, 240], np.array([[0, 1, 0], [1, 1, 1], [0, 0, 0]], dtype=np.uint8)),
        Tetromino(3, [0, 240, 0], np.array([[0, 1, 1], [1, 1, 0], [0, 0, 0]], dtype=np.uint8)),
        Tetromino(4, [240, 0, 0], np.array([[1, 1, 0], [0, 1, 1], [0, 0, 0]], dtype=np.uint8)),
        Tetromino(5, [0, 0, 240], np.array([[1, 0, 0], [1, 1, 1], [0, 0, 0]], dtype=np.uint8)),
        Tetromino(6, [240, 160, 0], np.array([[0, 0, 1], [1, 1, 1], [0, 0, 0]], dtype=np.uint8)),
    ]

    def __init__(self, render_mode=None, width=10, height=20, gravity=True,
                 actions_mapping=ActionsMapping(), rewards_mapping=RewardsMapping(),
                 queue=None, holder=None, randomizer=None, base_pixels=None, tetrominoes=None,
                 render_upscale: int = 10, *, num_envs: int = 1, device=None, queue_size: Optional[int] = None,
                 padding: Optional[int] = None, autoreset_mode: str = "next_step",
                 randomizer_mode: str = "philox", piece_sequences=None, env_id_offset: int = 0,
                 terminate_on_illegal_action: bool = True, report_invalid_actions: bool = False):
        if base_pixels is not None:
            # (the reference itself never assigns self.base_pixels when the argument is given, envs/tetris.py:113-115)
            raise NotImplementedError("custom base pixels are not supported: the empty / bedrock values 0 / 1 are built into the kernels")
        self.tetrominoes = None if tetrominoes is None else list(tetrominoes)
        if self.tetrominoes is not None and not 1 <= len(self.tetrominoes) <= 7:
            raise ValueError("a custom tetromino set holds 1..7 pieces")
        if padding not in (None, PADDING):
            raise ValueError("padding is derived from the tetromino set and must be 4")
        holder_size = int(getattr(holder, "size", holder)) if holder is not None else 1
        if not 1 <= holder_size <= 4:
            raise ValueError("holder size must be 1..4")
        if queue is not None and queue_size is None:
            queue_size = int(getattr(queue, "size", queue))
        if randomizer is None and getattr(queue, "randomizer", None) is not None:
            randomizer = queue.randomizer
        self.randomizer_kind = "bag"
        if randomizer is not None:
            name = randomizer if isinstance(randomizer, str) else (getattr(randomizer, "kind", None) or type(randomizer).__name__)
            if name in ("philox", "numpy", "sequence"):
                randomizer_mode = name
            elif name in ("true", "TrueRandomizer"):
                self.randomizer_kind = "true"
            elif name not in ("bag", "BagRandomizer"):
                raise NotImplementedError(f"randomizer {name!r}: only the reference's BagRandomizer / TrueRandomizer are built in")
        if not torch.cuda.is_available():
            raise RuntimeError("tetris_gymnasium_b200 needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("tetris_gymnasium_b200 runs on CUDA devices only")
        self.num_envs = int(num_envs)
        self.width, self.height = int(width), int(height)
        self.padding = PADDING
        self.width_padded = self.width + 2 * self.padding
        self.height_padded = self.height + self.padding
        self.queue_size = 4 if queue_size is None else int(queue_size)
        self.holder_size = holder_size   # > 1: a FIFO (TetrominoHolder.swap); the "holder" observation has the fixed shape (P, P * size)
        self.gravity_enabled = bool(gravity)
        self.actions, self.rewards = actions_mapping, rewards_mapping
        self.render_mode = render_mode
        self.render_scaling_factor = render_upscale
        self.autoreset_mode = autoreset_mode
        self.randomizer_mode = randomizer_mode
        self.env_id_offset = int(env_id_offset)
        self.observation_space, self.action_space = _spaces(self)
        self.single_observation_space, self.single_action_space = self.observation_space, self.action_space
        self.reward_range = (min(vars(self.rewards).values()), max(vars(self.rewards).values()))

        self._seq = None
        seq_len = 0
        if randomizer_mode == "sequence":
            if piece_sequences is None:
                raise ValueError("randomizer_mode='sequence' needs piece_sequences [num_envs, L]")
            self._seq = torch.as_tensor(np.asarray(piece_sequences) if not torch.is_tensor(piece_sequences) else piece_sequences)
            self._seq = self._seq.to(self.device, torch.uint8).contiguous()
            if self._seq.dim() == 1:
                self._seq = self._seq.unsqueeze(0).expand(self.num_envs, -1).contiguous()
            assert self._seq.shape[0] == self.num_envs
            seq_len = self._seq.shape[1]

        cfg = _lib.TgConfig()
        cfg.width, cfg.height, cfg.queue_size, cfg.gravity = self.width, self.height, self.queue_size, int(self.gravity_enabled)
        cfg.autoreset = _lib.AUTORESET[autoreset_mode]
        cfg.rng_mode = _lib.RNG[randomizer_mode]
        cfg.randomizer = _lib.RANDOMIZER[self.randomizer_kind]
        for i, f in enumerate(fields(ActionsMapping)):
            cfg.action_map[i] = int(getattr(self.actions, f.name))
        cfg.terminate_on_illegal = int(bool(terminate_on_illegal_action))
        cfg.reward_alife, cfg.reward_clear_line = float(self.rewards.alife), float(self.rewards.clear_line)
        cfg.reward_game_over, cfg.reward_invalid_action = float(self.rewards.game_over), float(self.rewards.invalid_action)
        cfg.seq_len, cfg.env_id_offset = seq_len, self.env_id_offset
        cfg.holder_size = self.holder_size
        if self.tetrominoes is not None:
            # Tetris(tetrominoes=[...]) (envs/tetris.py:117-132): matrices are used as binary masks, board values are index + 2
            cfg.n_pieces = len(self.tetrominoes)
            for i, t in enumerate(self.tetrominoes):
                m = np.asarray(t.matrix)
                if m.ndim != 2 or m.shape[0] != m.shape[1] or m.shape[0] > 4:
                    raise ValueError(f"tetromino {i}: the matrix must be square and at most 4 x 4")
                cfg.piece_n[i] = m.shape[0]
                for k, v in enumerate((m != 0).astype(np.uint8).reshape(-1)):
                    cfg.piece_matrix[i][k] = int(v)
                for k in range(3):
                    cfg.piece_color[i][k] = int(t.color_rgb[k])
        self._cfg = cfg
        self._L = _lib.load()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._L.tg_create(C.byref(cfg), self.device.index or 0, C.byref(h)))
        self._h = h
        lay = _lib.TgLayout()
        _lib.check(self._L.tg_get_layout(self._h, C.byref(lay)), self._h)
        self.layout = lay

        n, dev, u8 = self.num_envs, self.device, torch.uint8
        # state: caller-owned device memory (see tg_state in include/tetris_b200.h)
        self._hot = torch.zeros(n * lay.hot_stride, dtype=u8, device=dev)
        self._brd = torch.zeros(n * lay.board_stride, dtype=u8, device=dev)
        self._rng = torch.zeros(n * lay.rng_stride, dtype=u8, device=dev)
        # outputs (overwritten by every reset/step call)
        self._o_board = torch.empty((n, lay.height_padded, lay.width_padded), dtype=u8, device=dev)
        self._o_mask = torch.empty_like(self._o_board)
        self._o_holder = torch.empty((n, PADDING, PADDING * self.holder_size), dtype=u8, device=dev)
        self._o_queue = torch.empty((n, PADDING, PADDING * self.queue_size), dtype=u8, device=dev)
        self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._terminated = torch.zeros(n, dtype=u8, device=dev)
        self._truncated = torch.zeros(n, dtype=u8, device=dev)
        # the bool views the 5-tuple returns are made once (a tensor view costs more than a microsecond per call)
        self._terminated_b, self._truncated_b = self._terminated.view(torch.bool), self._truncated.view(torch.bool)
        self._lines = torch.zeros(n, dtype=torch.int32, device=dev)
        self._stats = torch.zeros(4, dtype=torch.float64, device=dev)
        self._seeded = False
        self._has_reset = False
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.report_invalid_actions = bool(report_invalid_actions)
        self._all_true = torch.ones(n, dtype=torch.bool, device=dev)
        self.emit_obs_dict = True   # wrappers that replace the observation (RgbObservation, FeatureVector) switch it off

    # ---- plumbing ---------------------------------------------------------------------------
    # The state / output buffers are allocated once, so the ctypes structs that carry their pointers are built once too
    # (for small batches the per-call Python overhead, not the kernel, is the step time).
    def _state(self):
        c = self.__dict__.get("_c_state")
        if c is None:
            c = self._c_state = _lib.TgState(self._hot.data_ptr(), self._brd.data_ptr(), self._rng.data_ptr(),
                                             self._seq.data_ptr() if self._seq is not None else None)
        return c

    def _obs_struct(self):
        c = self.__dict__.get("_c_obs")
        if c is None:
            c = self._c_obs = _lib.TgObs(self._o_board.data_ptr(), self._o_mask.data_ptr(), self._o_holder.data_ptr(), self._o_queue.data_ptr())
        return c

    def _out_struct(self):
        c = self.__dict__.get("_c_out")
        if c is None:
            c = self._c_out = _lib.TgStepOut(self._reward.data_ptr(), self._terminated.data_ptr(), self._truncated.data_ptr(), self._lines.data_ptr())
        return c

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _vector_info(self, info):
        """gymnasium's vector-env info format (SyncVectorEnv._add_info): every key `k` comes with a boolean mask `_k` saying which
        envs reported it -- here every env reports every key every step (examples/train_lin_grouped.py:316-322 reads
        infos["board"][0], infos["action_mask"][0])."""
        for k in list(info):
            if not k.startswith("_") and ("_" + k) not in info:
                info["_" + k] = self._all_true
        return info

    def _obs(self):
        return {"board": self._o_board, "active_tetromino_mask": self._o_mask, "holder": self._o_holder, "queue": self._o_queue}

    @property
    def unwrapped(self):
        return self

    def close(self):
        if getattr(self, "_h", None):
            self._L.tg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Tetris.reset (reference envs/tetris.py:274-307) -----------------------------------------
    def _per_env_seeds(self, seed):
        if torch.is_tensor(seed):
            seed = seed.cpu().numpy()
        if np.ndim(seed) == 0:
            return (np.uint64(int(seed)) + np.arange(self.num_envs, dtype=np.uint64) + np.uint64(self.env_id_offset)).astype(np.uint64)
        s = np.asarray(seed).astype(np.uint64)
        assert s.shape == (self.num_envs,)
        return s

    def _seed_numpy(self, seeds, mask=None):
        """Randomizer.reset (components/tetromino_randomizer.py:34-46): `default_rng(seed)` = PCG64(SeedSequence(seed)) for
        seed > 0; seed None / 0 keeps the running generator, which starts from OS entropy like `default_rng()`.  The
        SeedSequence hash and the PCG64 seeding run on the device (tg_seed_numpy_seeds): no per-env host work."""
        n = self.num_envs
        sel = np.ones(n, dtype=np.uint8) if mask is None else np.asarray(mask, dtype=np.uint8).copy()
        entropy = np.frombuffer(os.urandom(8 * n), dtype=np.uint64)
        if seeds is None:
            s = entropy
            if self._seeded:
                sel[:] = 0
        else:
            s = np.asarray(seeds, dtype=np.uint64).copy()
            unseeded = s == 0                                             # "if seed and seed > 0"
            s[unseeded] = entropy[unseeded]
            if self._seeded:
                sel[unseeded] = 0
        if not self._seeded:
            # first seeding: every env needs a valid generator, also the ones a partial reset leaves alone (an all-zero PCG64
            # state / increment is a degenerate stream)
            outside = sel == 0
            s = s.copy()
            s[outside] = entropy[outside]
            sel[:] = 1
        if not sel.any():
            return
        d_s = torch.from_numpy(s.view(np.int64)).to(self.device)
        d_m = torch.from_numpy(sel).to(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.tg_seed_numpy_seeds(self._h, self._state(), n, d_s.data_ptr(), d_m.data_ptr(), self._stream()), self._h)

    def reset(self, *, seed=None, options: "dict[str, Any] | None" = None):
        """Reset all envs (or `options["reset_mask"]`).  `seed`: int (env i gets seed + i, like
        gymnasium's vector envs) or one seed per env.  Returns (obs dict, {"lines_cleared": 0})."""
        mask = None
        if options and options.get("reset_mask") is not None:
            mask = torch.as_tensor(options["reset_mask"]).to(self.device).to(torch.uint8).contiguous()
        d_seeds = None
        if self.randomizer_mode == "numpy":
            if seed is not None or not self._seeded:
                seeds = None if seed is None else self._per_env_seeds(seed)
                self._seed_numpy(seeds, None if mask is None else mask.cpu().numpy())
                self._seeded = True
        elif self.randomizer_mode == "philox":
            if seed is not None or not self._seeded:
                seeds = self._per_env_seeds(0x5EED if seed is None else seed)
                d_seeds = torch.from_numpy(seeds.view(np.int64)).to(self.device)
                self._seeded = True
        with torch.cuda.device(self.device):
            _lib.check(self._L.tg_reset(self._h, self._state(), self.num_envs,
                                        d_seeds.data_ptr() if d_seeds is not None else None,
                                        mask.data_ptr() if mask is not None else None,
                                        self._obs_struct(), self._stream()), self._h)
        self._has_reset = True
        self._lines.zero_()
        return self._obs(), self._vector_info({"lines_cleared": self._lines})

    # ---- Tetris.step (reference envs/tetris.py:203-272) -------------------------------------------
    def _actions(self, actions):
        if torch.is_tensor(actions) and actions.dtype == torch.int32 and actions.device == self.device and actions.is_contiguous() \
                and actions.shape == (self.num_envs,):
            return actions                                   # the usual case: no conversion kernels, no checks
        a = actions if torch.is_tensor(actions) else torch.as_tensor(np.asarray(actions))
        a = a.to(device=self.device, dtype=torch.int32, non_blocking=True).contiguous()
        if a.dim() == 0:
            a = a.reshape(1)
        assert a.shape == (self.num_envs,), f"actions must have shape ({self.num_envs},)"
        return a

    def step(self, actions):
        """One step of every env.  Returns (obs, reward f32[n], terminated bool[n], truncated bool[n],
        {"lines_cleared": i32[n]}) -- tensors are the env's output buffers, overwritten by the next call."""
        a = self._actions(actions)
        obs = self._obs_struct() if self.emit_obs_dict else _NO_OBS
        # tg_step selects the env's device itself; the guard only keeps torch's notion of the current device in step
        if torch.cuda.current_device() == self._dev_index:
            rc = self._L.tg_step(self._h, self._state(), self.num_envs, a.data_ptr(), obs, self._out_struct(), self._stats.data_ptr(), self._stream())
        else:
            with torch.cuda.device(self.device):
                rc = self._L.tg_step(self._h, self._state(), self.num_envs, a.data_ptr(), obs, self._out_struct(), self._stats.data_ptr(), self._stream())
        if rc:
            _lib.check(rc, self._h)
        info = {"lines_cleared": self._lines, "_lines_cleared": self._all_true}
        if self.report_invalid_actions:
            # the reference asserts on an action outside the action space (envs/tetris.py:215); the batched env treats it as the
            # unmatched elif chain (no move) and reports it per env instead of aborting the whole batch
            info["invalid_action"] = (a < 0) | (a >= 8)
            info["_invalid_action"] = self._all_true
        return (self._obs(), self._reward, self._terminated_b, self._truncated_b, info)

    def step_n(self, actions, keep_all: bool = True):
        """K consecutive steps in one native call (tg_step_n): `actions` int32 [K, num_envs] on the device.
        keep_all=True returns rollout storage with a leading step axis -- obs dict arrays [K, n, ...], reward / terminated /
        truncated [K, n], info["lines_cleared"] [K, n] (buffers reused by the next call of the same K); keep_all=False keeps
        only the last step's observation and 5-tuple in the env's usual output buffers.  Batches whose records fit in shared
        memory run as ONE persistent launch with the state resident on chip for all K steps (small batches are otherwise bound
        by launch / call latency, not by work); results are identical to K step() calls."""
        a = actions if torch.is_tensor(actions) else torch.as_tensor(np.asarray(actions))
        a = a.to(device=self.device, dtype=torch.int32).contiguous()
        assert a.dim() == 2 and a.shape[1] == self.num_envs, f"actions must have shape (K, {self.num_envs})"
        K, n, lay, u8 = int(a.shape[0]), self.num_envs, self.layout, torch.uint8
        if keep_all:
            st = self.__dict__.get("_stepn")
            if st is None or st["K"] != K:
                dev = self.device
                st = self._stepn = {
                    "K": K,
                    "board": torch.empty((K, n, lay.height_padded, lay.width_padded), dtype=u8, device=dev),
                    "mask": torch.empty((K, n, lay.height_padded, lay.width_padded), dtype=u8, device=dev),
                    "holder": torch.empty((K, n, PADDING, PADDING * self.holder_size), dtype=u8, device=dev),
                    "queue": torch.empty((K, n, PADDING, PADDING * self.queue_size), dtype=u8, device=dev),
                    "reward": torch.empty((K, n), dtype=torch.float32, device=dev), "terminated": torch.empty((K, n), dtype=u8, device=dev),
                    "truncated": torch.empty((K, n), dtype=u8, device=dev), "lines": torch.empty((K, n), dtype=torch.int32, device=dev)}
                st["c_obs"] = _lib.TgObs(st["board"].data_ptr(), st["mask"].data_ptr(), st["holder"].data_ptr(), st["queue"].data_ptr())
                st["c_out"] = _lib.TgStepOut(st["reward"].data_ptr(), st["terminated"].data_ptr(), st["truncated"].data_ptr(), st["lines"].data_ptr())
            obs, out, stride = st["c_obs"], st["c_out"], n
        else:
            obs, out, stride = self._obs_struct(), self._out_struct(), 0
        with torch.cuda.device(self.device):
            _lib.check(self._L.tg_step_n(self._h, self._state(), n, K, a.data_ptr(), obs, stride, out, stride, self._stats.data_ptr(),
                                         self._stream()), self._h)
        if keep_all:
            return ({"board": st["board"], "active_tetromino_mask": st["mask"], "holder": st["holder"], "queue": st["queue"]},
                    st["reward"], st["terminated"].view(torch.bool), st["truncated"].view(torch.bool), {"lines_cleared": st["lines"]})
        return (self._obs(), self._reward, self._terminated_b, self._truncated_b, {"lines_cleared": self._lines})

    def step_host(self, actions: np.ndarray, out: "dict[str, np.ndarray] | None" = None, mode: str = "compact"):
        """Same step with HOST arrays in and out (tg_step_host) -- the reference's own calling convention.
        mode="compact" (default): the step runs without the dict, the packed state records cross PCIe and the dict is rebuilt
        in `out` by the library's host threads; mode="dma": the dict is written on the device and DMA-copied.  Both give
        identical arrays.  `out` may hold preallocated (ideally pinned, `alloc_host_buffers`) numpy arrays; returns them."""
        n = self.num_envs
        a = actions if (isinstance(actions, np.ndarray) and actions.dtype == np.int32 and actions.flags.c_contiguous) \
            else np.ascontiguousarray(actions, dtype=np.int32)
        assert a.shape == (n,), f"actions must have shape ({n},)"
        if out is None:
            out = self.alloc_host_buffers(pinned=False)
        key = (id(out), mode)
        cached = self.__dict__.get("_c_host")
        if cached is None or cached[0] != key:
            ho = _lib.TgObs(out["board"].ctypes.data, out["active_tetromino_mask"].ctypes.data, out["holder"].ctypes.data, out["queue"].ctypes.data)
            so = _lib.TgStepOut(out["reward"].ctypes.data, out["terminated"].ctypes.data, out["truncated"].ctypes.data, out["lines_cleared"].ctypes.data)
            cached = self._c_host = (key, ho, so, out)      # keeps `out` alive while its pointers are cached
        with torch.cuda.device(self.device):
            _lib.check(self._L.tg_step_host(self._h, self._state(), n, a.ctypes.data, cached[1], cached[2],
                                            _lib.HOST_MODE[mode], self._stream()), self._h)
        return out

    def set_host_threads(self, threads: int = 0):
        """Host threads of step_host(mode="compact"); 0 = cores of this process / LOCAL_WORLD_SIZE."""
        _lib.check(self._L.tg_set_host_threads(self._h, int(threads)), self._h)

    def host_stats(self):
        """Timing of the last step_host call: seconds total / waiting for the device / expanding, and the chunk count."""
        v = (C.c_double * 4)()
        _lib.check(self._L.tg_host_stats(self._h, v), self._h)
        return {"total_s": v[0], "wait_s": v[1], "expand_s": v[2], "chunks": int(v[3])}

    def alloc_host_buffers(self, pinned=True):
        n, lay = self.num_envs, self.layout
        def mk(shape, dt):
            t = torch.empty(shape, dtype=dt, pin_memory=pinned)
            return t.numpy()
        return {
            "board": mk((n, lay.height_padded, lay.width_padded), torch.uint8),
            "active_tetromino_mask": mk((n, lay.height_padded, lay.width_padded), torch.uint8),
            "holder": mk((n, PADDING, PADDING * self.holder_size), torch.uint8),
            "queue": mk((n, PADDING, PADDING * self.queue_size), torch.uint8),
            "reward": mk((n,), torch.float32), "terminated": mk((n,), torch.uint8),
            "truncated": mk((n,), torch.uint8), "lines_cleared": mk((n,), torch.int32),
        }

    # ---- fused heuristic rollout (tg_rollout; BASELINE config 4) ----------------------------------------
    def rollout(self, weights, k_steps: int, trace: bool = False):
        """K grouped-placement steps of the integer linear policy
        score = w0*sum(heights) + w1*lines + w2*holes + w3*bumpiness (lowest index among the legal maxima),
        fused in one kernel launch with boards resident on chip.  State advances in place; episode statistics
        accumulate into `episode_stats()`.  With trace=True returns the action chosen at the last step."""
        w = (C.c_int32 * 4)(*[int(v) for v in weights])
        last = None
        if trace:
            last = torch.full((self.num_envs,), -1, dtype=torch.int32, device=self.device)
        _lib.check(self._L.tg_debug_set_rollout_trace(self._h, last.data_ptr() if last is not None else None), self._h)
        with torch.cuda.device(self.device):
            _lib.check(self._L.tg_rollout(self._h, self._state(), self.num_envs, w, int(k_steps), self._stats.data_ptr(),
                                          self._stream()), self._h)
        return last

    # ---- episode statistics (RecordEpisodeStatistics-style, accumulated on device) ----------------
    def episode_stats(self, reset=False):
        s = self._stats.clone()
        if reset:
            self._stats.zero_()
        return {"episodes": s[0], "sum_return": s[1], "sum_length": s[2], "sum_lines": s[3]}

    # ---- state access (reference get_state/set_state :681-708 and the tests' env.unwrapped pokes) --
    def get_state(self):
        """Unpacked state: board u8[n,Hp,Wp] (locked cells), x, y, piece, rotation, holder, ... tensors."""
        n, lay = self.num_envs, self.layout
        board = torch.empty((n, lay.height_padded, lay.width_padded), dtype=torch.uint8, device=self.device)
        Q, S = self.queue_size, self.holder_size
        sc = torch.empty((n, _lib.TG_SCALARS + Q + (2 * S if S > 1 else 0)), dtype=torch.int32, device=self.device)
        _lib.check(self._L.tg_get_state(self._h, self._state(), n, board.data_ptr(), sc.data_ptr(), self._stream()), self._h)
        st = {"board": board, "x": sc[:, 0], "y": sc[:, 1], "piece": sc[:, 2], "rotation": sc[:, 3],
              "holder_piece": sc[:, 4], "holder_rotation": sc[:, 5], "has_swapped": sc[:, 6], "game_over": sc[:, 7],
              "queue": sc[:, 8:8 + Q], "_scalars": sc,
              "_raw": (self._hot.clone(), self._brd.clone(), self._rng.clone())}
        if S > 1:   # FIFO holder: number of held pieces, then (piece, rotation) per slot, oldest first, -1 = empty slot
            del st["holder_piece"], st["holder_rotation"]
            st["holder_count"] = sc[:, 4]
            st["holder_pieces"] = sc[:, 8 + Q::2]
            st["holder_rotations"] = sc[:, 9 + Q::2]
        return st

    def set_state(self, state=None, *, board=None, env_mask=None, **scalars):
        """Restore a `get_state()` snapshot, or poke fields: set_state(board=..., x=..., piece=..., ...)."""
        n = self.num_envs
        if state is not None and "_raw" in state and board is None and not scalars:
            self._hot.copy_(state["_raw"][0]); self._brd.copy_(state["_raw"][1]); self._rng.copy_(state["_raw"][2])
            return
        cur = self.get_state() if state is None else state
        sc = cur["_scalars"].clone()
        names = {"x": 0, "y": 1, "piece": 2, "rotation": 3, "holder_piece": 4, "holder_rotation": 5, "has_swapped": 6, "game_over": 7}
        Q = self.queue_size
        for k, v in scalars.items():
            if k == "queue":
                sc[:, 8:8 + Q] = torch.as_tensor(v, device=self.device).to(torch.int32)
            elif k == "holder_pieces":
                sc[:, 8 + Q::2] = torch.as_tensor(v, device=self.device).to(torch.int32)
            elif k == "holder_rotations":
                sc[:, 9 + Q::2] = torch.as_tensor(v, device=self.device).to(torch.int32)
            else:
                sc[:, names[k]] = torch.as_tensor(v, device=self.device).to(torch.int32)
        b = None
        if board is not None:
            b = torch.as_tensor(np.asarray(board) if not torch.is_tensor(board) else board).to(self.device, torch.uint8)
            if b.dim() == 2:
                b = b.unsqueeze(0).expand(n, -1, -1)
            b = b.contiguous()
        elif state is not None:
            b = state["board"].contiguous()
        m = None if env_mask is None else torch.as_tensor(env_mask).to(self.device).to(torch.uint8).contiguous()
        _lib.check(self._L.tg_set_state(self._h, self._state(), n, b.data_ptr() if b is not None else None,
                                        sc.data_ptr(), m.data_ptr() if m is not None else None, self._stream()), self._h)

    def observe(self):
        """Re-emit the observation dict of the current state (Tetris._get_obs) without stepping."""
        with torch.cuda.device(self.device):
            zero = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
            _lib.check(self._L.tg_reset(self._h, self._state(), self.num_envs, None, zero.data_ptr(),
                                        self._obs_struct(), self._stream()), self._h)
        return self._obs()

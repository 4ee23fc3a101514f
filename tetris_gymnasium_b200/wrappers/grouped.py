"""GroupedActionsObservations -- mirrors tetris_gymnasium/wrappers/grouped.py on the batched CUDA env.

Action space Discrete(4*W): action = column * 4 + rotations (reference :78-99).  Each observation
holds one entry per placement: the resulting board (no observation wrappers) or its feature vector
(observation_wrappers=[FeatureVectorObservation(env)]), plus info["action_mask"], info["board"],
info["lines_cleared"] like the reference (:209-294).  Enumeration and execution run in
tg_grouped_observe / tg_grouped_step.
"""
import numpy as np
import torch

from .. import _lib
from .observation import FeatureVectorObservation


def u_high(w):
    u = w.unwrapped
    return u.height * u.width


_NO_OBS = _lib.TgObs(None, None, None, None)


class GroupedActionsObservations:
    """Constructor as in the reference (wrappers/grouped.py:44-49).  Keyword-only extras:

      mask_dtype   dtype of `legal_actions_mask` / info["action_mask"]: torch.uint8 (default, what the kernels write) or
                   torch.float64 (the reference's `np.ones(4 * W)` dtype) / any float dtype for drop-in trainers
      obs_dtype    None (uint8, what the kernels write) or a float dtype: the observation is cast on return (the reference
                   returns float arrays after an illegal action, `np.ones_like(obs) * observation_space.high`)

    observation_wrappers: None / [] = placement board images; [FeatureVectorObservation(env)] = the fused feature kernel; any
    other list = the board images are produced natively and each wrapper's `observation()` is applied in turn to the batched
    dict {"board": u8[n * 4W, Hp, Wp], "active_tetromino_mask", "holder", "queue"} (leading axis = env x placement), the
    reference's per-placement loop (:184-204) as one batched call per wrapper; the result is reshaped to [n, 4W, ...]."""

    def __init__(self, env, observation_wrappers=None, terminate_on_illegal_action: bool = True, *, mask_dtype=torch.uint8, obs_dtype=None):
        self.env = env
        u = env.unwrapped
        # a wrapper option in the reference; the native handle carries it, so the wrapper sets it there
        _lib.check(u._L.tg_set_option(u._h, _lib.TG_OPT_TERMINATE_ON_ILLEGAL, int(bool(terminate_on_illegal_action))), u._h)
        self.observation_wrappers = observation_wrappers
        self.terminate_on_illegal_action = terminate_on_illegal_action
        self.mask_dtype, self.obs_dtype = mask_dtype, obs_dtype
        self._featw, self._generic = None, None
        if observation_wrappers:
            if len(observation_wrappers) == 1 and isinstance(observation_wrappers[0], FeatureVectorObservation):
                self._featw = observation_wrappers[0]
            else:
                self._generic = list(observation_wrappers)
        n, A, F = u.num_envs, u.layout.n_placements, u.layout.n_features
        from ..envs.tetris import _Space
        self.action_space = _Space(n=A, dtype=np.int64)
        dev = u.device
        self._legal = torch.ones((n, A), dtype=torch.uint8, device=dev)
        if self._featw is not None:
            self._feats = torch.empty((n, A, F), dtype=torch.uint8, device=dev)
            self._info_board = torch.empty((n, F), dtype=torch.uint8, device=dev)
            self._boards = None
            single = self._featw.observation_space.shape
        else:
            self._feats, self._info_board = None, None
            self._boards = torch.empty((n, A, u.height_padded, u.width_padded), dtype=torch.uint8, device=dev)
            single = (u.height_padded, u.width_padded)
        self.observation_space = _Space((A,) + tuple(single), np.float32, 0, u.height * u.width)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def legal_actions_mask(self):
        return self._legal if self.mask_dtype == torch.uint8 else self._legal.to(self.mask_dtype)

    def encode_action(self, x, r):
        return x * 4 + r

    def decode_action(self, action):
        return action // 4, action % 4

    def _ptr(self, t):
        return None if t is None else t.data_ptr()

    def _result(self):
        if self._featw is not None:
            out = self._featw.select(self._feats)
        elif self._generic is not None:
            u = self.unwrapped
            n, A = u.num_envs, u.layout.n_placements
            rep = lambda t: t.unsqueeze(1).expand((n, A) + tuple(t.shape[1:])).reshape((n * A,) + tuple(t.shape[1:]))  # noqa: E731
            obs = {"board": self._boards.reshape(n * A, u.height_padded, u.width_padded), "active_tetromino_mask": rep(u._o_mask),
                   "holder": rep(u._o_holder), "queue": rep(u._o_queue)}
            for w in self._generic:
                obs = w.observation(obs)
            out = obs.reshape((n, A) + tuple(obs.shape[1:])) if torch.is_tensor(obs) else obs
        else:
            out = self._boards
        if self.obs_dtype is not None and torch.is_tensor(out):
            out = out.to(self.obs_dtype)
            if self.terminate_on_illegal_action and u_high(self) > 255 and self._generic is None:
                # the uint8 kernels saturate the reference's "illegal action" fill (observation_space.high = H * W) at 255; a float
                # observation carries the true value: an env whose episode an illegal action just ended reads 255 everywhere
                u = self.unwrapped
                ended = u._terminated.view(torch.bool) & (u._reward == float(u.rewards.invalid_action))
                flat = out.reshape(out.shape[0], -1)
                ended = ended & (flat == 255).all(dim=1)
                out = torch.where(ended.reshape((-1,) + (1,) * (out.dim() - 1)), torch.full_like(out, float(u_high(self))), out)
        return out

    def observation(self, observation=None):
        """Enumerate all 4*W placements of the current state (reference :124-207)."""
        u = self.unwrapped
        with torch.cuda.device(u.device):
            _lib.check(u._L.tg_grouped_observe(u._h, u._state(), u.num_envs, self._ptr(self._feats), self._ptr(self._boards),
                                               self._legal.data_ptr(), u._stream()), u._h)
        return self._result()

    def _info(self, with_board=True):
        u = self.unwrapped
        info = {"action_mask": self.legal_actions_mask, "lines_cleared": u._lines}
        if with_board:
            info["board"] = self._featw.select(self._info_board) if self._featw is not None else u._obs()
        return u._vector_info(info)

    def reset(self, *, seed=None, options=None):
        u = self.unwrapped
        obs, _ = self.env.reset(seed=seed, options=options)
        if self._featw is not None:
            with torch.cuda.device(u.device):
                _lib.check(u._L.tg_features(u._h, u._state(), u.num_envs, self._info_board.data_ptr(), u._stream()), u._h)
        return self.observation(obs), self._info()

    def step(self, action):
        u = self.unwrapped
        a = u._actions(action)
        want_dict = self._featw is None   # (generic wrappers read the base dict's mask / holder / queue)
        obs = u._obs_struct() if want_dict else _NO_OBS
        c = self.__dict__.get("_c_ptrs")
        if c is None:      # the output buffers are allocated once: so are their pointers
            c = self._c_ptrs = (self._legal.data_ptr(), self._ptr(self._feats), self._ptr(self._boards), self._ptr(self._info_board))
        args = (u._h, u._state(), u.num_envs, a.data_ptr(), c[0], c[1], c[2], c[3], obs, u._out_struct(), u._stats.data_ptr(), u._stream())
        if torch.cuda.current_device() == u._dev_index:
            rc = u._L.tg_grouped_step(*args)
        else:
            with torch.cuda.device(u.device):
                rc = u._L.tg_grouped_step(*args)
        if rc:
            _lib.check(rc, u._h)
        return (self._result(), u._reward, u._terminated_b, u._truncated_b, self._info())

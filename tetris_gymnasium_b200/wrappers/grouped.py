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

_NO_OBS = _lib.TgObs(None, None, None, None)


class GroupedActionsObservations:
    def __init__(self, env, observation_wrappers=None, terminate_on_illegal_action: bool = True):
        self.env = env
        u = env.unwrapped
        if bool(terminate_on_illegal_action) != bool(u._cfg.terminate_on_illegal):
            raise ValueError("pass terminate_on_illegal_action to the Tetris constructor as well "
                             "(it is part of the native env config)")
        self.observation_wrappers = observation_wrappers
        self.terminate_on_illegal_action = terminate_on_illegal_action
        self._featw = None
        if observation_wrappers:
            if len(observation_wrappers) != 1 or not isinstance(observation_wrappers[0], FeatureVectorObservation):
                raise NotImplementedError("only [FeatureVectorObservation] is supported as observation_wrappers")
            self._featw = observation_wrappers[0]
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
        return self._legal

    def encode_action(self, x, r):
        return x * 4 + r

    def decode_action(self, action):
        return action // 4, action % 4

    def _ptr(self, t):
        return None if t is None else t.data_ptr()

    def _result(self):
        if self._featw is not None:
            return self._featw.select(self._feats)
        return self._boards

    def observation(self, observation=None):
        """Enumerate all 4*W placements of the current state (reference :124-207)."""
        u = self.unwrapped
        with torch.cuda.device(u.device):
            _lib.check(u._L.tg_grouped_observe(u._h, u._state(), u.num_envs, self._ptr(self._feats), self._ptr(self._boards),
                                               self._legal.data_ptr(), u._stream()), u._h)
        return self._result()

    def _info(self, with_board=True):
        u = self.unwrapped
        info = {"action_mask": self._legal, "lines_cleared": u._lines}
        if with_board:
            info["board"] = self._featw.select(self._info_board) if self._featw is not None else u._obs()
        return info

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
        want_dict = self._featw is None
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
        return (self._result(), u._reward, u._terminated.view(torch.bool), u._truncated.view(torch.bool), self._info())

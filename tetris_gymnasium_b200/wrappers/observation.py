"""Observation wrappers -- mirrors tetris_gymnasium/wrappers/observation.py on the batched CUDA env.

RgbObservation (reference :11-115)            -> tg_render_rgb
FeatureVectorObservation (reference :118-278) -> tg_features
Both return torch CUDA tensors with a leading env axis.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib


class _ObsWrapper:
    def __init__(self, env, keep_obs_dict=False):
        self.env = env
        self.action_space = env.action_space
        # the wrapper replaces the observation: the base env need not write the dict on step()
        if not keep_obs_dict:
            env.unwrapped.emit_obs_dict = False

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), reward, terminated, truncated, info

    def close(self):
        return self.env.close()


class RgbObservation(_ObsWrapper):
    """Board on the left, queue top right, holder bottom right, as one RGB image
    u8[n, H_pad, W_pad + max(queue, holder) * P, 3] (reference wrappers/observation.py:38-74)."""

    def __init__(self, env, keep_obs_dict=False):
        super().__init__(env, keep_obs_dict)
        u = env.unwrapped
        self.shape = (u.height_padded, u.layout.rgb_width, 3)
        self._img = torch.empty((u.num_envs,) + self.shape, dtype=torch.uint8, device=u.device)
        from ..envs.tetris import _Space
        self.observation_space = _Space(self.shape, np.uint8, 0, 255)

    def observation(self, observation=None):
        u = self.env.unwrapped
        with torch.cuda.device(u.device):
            _lib.check(u._L.tg_render_rgb(u._h, u._state(), u.num_envs, self._img.data_ptr(), u._stream()), u._h)
        return self._img


class FeatureVectorObservation(_ObsWrapper):
    """heights(W), max height, holes, bumpiness as u8[n, W+3] (reference wrappers/observation.py:238-278),
    including the reference's row-0/1 zeroing through integer indexing and the uint8 wrap (Q4)."""

    def __init__(self, env, report_height=True, report_max_height=True, report_holes=True, report_bumpiness=True,
                 keep_obs_dict=True):
        super().__init__(env, keep_obs_dict)
        u = env.unwrapped
        self.report_height, self.report_max_height = report_height, report_max_height
        self.report_holes, self.report_bumpiness = report_holes, report_bumpiness
        W = u.width
        cols = (list(range(W)) if report_height else []) + ([W] if report_max_height else []) + \
               ([W + 1] if report_holes else []) + ([W + 2] if report_bumpiness else [])
        self._all = len(cols) == W + 3
        self._cols = torch.tensor(cols, dtype=torch.long, device=u.device)
        self._feats = torch.empty((u.num_envs, W + 3), dtype=torch.uint8, device=u.device)
        from ..envs.tetris import _Space
        self.observation_space = _Space((len(cols),), np.uint8, 0, 7)

    def select(self, feats):
        """Keep the reported features (last axis)."""
        return feats if self._all else feats.index_select(-1, self._cols)

    def observation(self, observation=None):
        u = self.env.unwrapped
        with torch.cuda.device(u.device):
            _lib.check(u._L.tg_features(u._h, u._state(), u.num_envs, self._feats.data_ptr(), u._stream()), u._h)
        return self.select(self._feats)

"""Observation wrappers -- mirrors tetris_gymnasium/wrappers/observation.py on the batched CUDA env.

RgbObservation (reference :11-115)            -> tg_render_rgb
FeatureVectorObservation (reference :118-278) -> tg_features
CnnObservation (examples/train_cnn.py:127-147: RgbObservation -> ResizeObservation(84, 84) -> GrayscaleObservation ->
FrameStackObservation(4) [+ ClipRewardEnv])   -> tg_cnn_observe, one fused kernel
All return torch CUDA tensors with a leading env axis.
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


class CnnObservation(_ObsWrapper):
    """The image pipeline of the reference's CNN trainer (examples/train_cnn.py:127-147) as one fused CUDA kernel:
    RgbObservation -> gym.wrappers.ResizeObservation(shape) (cv2 INTER_AREA) -> GrayscaleObservation ->
    FrameStackObservation(stack_size) (reset frame repeated at reset), optionally ClipRewardEnv (sign of the reward).

    Observation: u8[n, stack_size, H, W], oldest frame first -- a VIEW into a sliding window of `stack_size + window`
    frames per env (each step writes one new 7 KB frame; every `window` steps the last stack_size-1 frames are moved to the
    front), valid until the next step()/reset() call."""

    def __init__(self, env, shape=(84, 84), stack_size=4, window=28, clip_reward=False):
        super().__init__(env, keep_obs_dict=False)
        u = env.unwrapped
        self.shape, self.stack_size, self.clip_reward = (int(shape[0]), int(shape[1])), int(stack_size), bool(clip_reward)
        self._slots = self.stack_size + int(window)
        self._fb = self.shape[0] * self.shape[1]
        self._buf = torch.zeros((u.num_envs, self._slots) + self.shape, dtype=torch.uint8, device=u.device)
        self._cur = self.stack_size - 1
        self._prev_term = torch.zeros(u.num_envs, dtype=torch.uint8, device=u.device)
        self._ones = torch.ones(u.num_envs, dtype=torch.uint8, device=u.device)
        from ..envs.tetris import _Space
        self.observation_space = _Space((self.stack_size,) + self.shape, np.uint8, 0, 255)

    def _emit(self, fill_mask):
        u = self.env.unwrapped
        with torch.cuda.device(u.device):
            _lib.check(u._L.tg_cnn_observe(u._h, u._state(), u.num_envs, self.shape[0], self.shape[1],
                                           self._buf.data_ptr() + self._cur * self._fb, self._slots * self._fb,
                                           fill_mask.data_ptr() if fill_mask is not None else None, self.stack_size - 1,
                                           u._stream()), u._h)
        return self._buf[:, self._cur - self.stack_size + 1: self._cur + 1]

    def observation(self, observation=None):
        """Current stack window (re-rendering the newest frame from the current state)."""
        return self._emit(None)

    def reset(self, *, seed=None, options=None):
        _, info = self.env.reset(seed=seed, options=options)
        mask = self._ones
        if options and options.get("reset_mask") is not None:
            mask = torch.as_tensor(options["reset_mask"]).to(self._buf.device).to(torch.uint8).contiguous()
        self._prev_term.zero_()
        return self._emit(mask), info

    def step(self, action):
        u = self.env.unwrapped
        _, reward, terminated, truncated, info = self.env.step(action)
        self._cur += 1
        if self._cur == self._slots:   # slide the window back to the front
            k = self.stack_size - 1
            self._buf[:, :k].copy_(self._buf[:, self._slots - k:].clone())
            self._cur = k
        if u.autoreset_mode == "next_step":
            fill = self._prev_term.clone()     # envs that were reset by this call (their action was ignored)
            self._prev_term.copy_(terminated.view(torch.uint8))
        elif u.autoreset_mode == "same_step":
            fill = terminated.view(torch.uint8)
        else:
            fill = None
        obs = self._emit(fill)
        if self.clip_reward:
            reward = torch.sign(reward)    # stable_baselines3 ClipRewardEnv (examples/train_cnn.py:138)
        return obs, reward, terminated, truncated, info

"""Trainer-side bookkeeping wrappers of the reference's examples, on device tensors.

RecordEpisodeStatistics mirrors `gym.wrappers.RecordEpisodeStatistics` as the examples use it inside a
`SyncVectorEnv` (examples/train_lin_grouped.py:148, examples/train_cnn.py:135), in the vector-env info format
(`info["episode"] = {"r", "l", "t"}` with the `info["_episode"]` mask; gymnasium 1.x wrappers/vector/common.py -- third-party,
restated from its published behaviour).  Everything is a handful of elementwise torch ops on [n] tensors: plumbing, not a kernel.
The library also accumulates whole-run totals on device (`env.unwrapped.episode_stats()`, tg_stats)."""
import time

import torch


class RecordEpisodeStatistics:
    def __init__(self, env):
        self.env = env
        u = env.unwrapped
        n, dev = u.num_envs, u.device
        self.episode_returns = torch.zeros(n, dtype=torch.float32, device=dev)
        self.episode_lengths = torch.zeros(n, dtype=torch.int32, device=dev)
        self.prev_dones = torch.zeros(n, dtype=torch.bool, device=dev)
        self.episode_count = 0
        self._t0 = time.perf_counter()
        self.action_space = getattr(env, "action_space", None)
        self.observation_space = getattr(env, "observation_space", None)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, *, seed=None, options=None):
        out = self.env.reset(seed=seed, options=options)
        mask = None
        if options and options.get("reset_mask") is not None:
            mask = torch.as_tensor(options["reset_mask"]).to(self.episode_returns.device).to(torch.bool)
        if mask is None:
            self.episode_returns.zero_(); self.episode_lengths.zero_(); self.prev_dones.zero_()
        else:   # a partial reset only restarts the masked envs' episodes
            self.episode_returns.masked_fill_(mask, 0.0); self.episode_lengths.masked_fill_(mask, 0); self.prev_dones.masked_fill_(mask, False)
        self._t0 = time.perf_counter()
        return out

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        next_step = getattr(self.unwrapped, "autoreset_mode", "next_step") == "next_step"
        if next_step:
            # envs that finished on the previous step are being reset by this call (NEXT_STEP autoreset: their action is ignored,
            # reward 0): the call is not a step of any episode
            self.episode_returns.masked_fill_(self.prev_dones, 0.0)
            self.episode_lengths.masked_fill_(self.prev_dones, 0)
            live = ~self.prev_dones
            self.episode_returns += reward * live
            self.episode_lengths += live.to(torch.int32)
        else:
            # SAME_STEP / disabled autoreset: every call is a real step; the accumulators restart right after an episode is reported
            self.episode_returns += reward
            self.episode_lengths += 1
        dones = terminated | truncated
        info = dict(info)
        zero = torch.zeros((), device=reward.device)
        info["episode"] = {"r": torch.where(dones, self.episode_returns, zero),
                           "l": torch.where(dones, self.episode_lengths, zero.to(torch.int32)),
                           "t": torch.where(dones, torch.full_like(self.episode_returns, time.perf_counter() - self._t0), zero)}
        info["_episode"] = dones
        if next_step:
            self.prev_dones = dones.clone()
        else:
            self.episode_returns = self.episode_returns.masked_fill(dones, 0.0)     # (new tensors: the reported ones stay valid)
            self.episode_lengths = self.episode_lengths.masked_fill(dones, 0)
        return obs, reward, terminated, truncated, info

    def close(self):
        return self.env.close()

"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

from oracle.tetris_oracle import OracleEnv


def np_(t):
    return t.detach().cpu().numpy()


class OracleBatch:
    """n oracle envs mirroring a batched CUDA env (autoreset handled like gymnasium NEXT_STEP)."""

    def __init__(self, n, seqs=None, seeds=None, **kw):
        self.envs = [OracleEnv(**kw) for _ in range(n)]
        self.n = n
        self.pending = np.zeros(n, bool)
        for i, e in enumerate(self.envs):
            if seqs is not None:
                e.set_sequence(seqs[i])
            elif seeds is not None:
                e.seed_numpy(int(seeds[i]))

    def reset(self):
        obs = [e.reset()[0] for e in self.envs]
        self.pending[:] = False
        return {k: np.stack([o[k] for o in obs]) for k in obs[0]}

    def step(self, actions, autoreset="next_step"):
        obs, rew, term, lines = [], [], [], []
        for i, e in enumerate(self.envs):
            if autoreset == "next_step" and self.pending[i]:
                o, _ = e.reset()
                r, t, l = 0.0, False, 0
            else:
                o, r, t, _, info = e.step(int(actions[i]))
                l = info["lines_cleared"]
                if autoreset == "same_step" and t:
                    o, _ = e.reset()
            obs.append(o); rew.append(r); term.append(t); lines.append(l)
        self.pending = np.array(term) if autoreset == "next_step" else np.zeros(self.n, bool)
        return ({k: np.stack([o[k] for o in obs]) for k in obs[0]}, np.array(rew, np.float32),
                np.array(term), np.array(lines, np.int32))


def assert_obs_equal(got, want, ctx=""):
    for k in ("board", "active_tetromino_mask", "holder", "queue"):
        g = np_(got[k])
        assert g.dtype == np.uint8 and g.shape == want[k].shape, (ctx, k, g.shape, want[k].shape)
        if not np.array_equal(g, want[k]):
            bad = np.flatnonzero((g != want[k]).reshape(len(g), -1).any(1))
            raise AssertionError(f"{ctx}: obs[{k}] differs for envs {bad[:8]} (of {len(bad)})\n got:\n{g[bad[0]]}\n want:\n{want[k][bad[0]]}")

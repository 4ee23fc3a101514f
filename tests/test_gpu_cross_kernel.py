"""Large-batch cross-checks between independent CUDA implementations of the same path.

The oracle comparisons (test_gpu_grouped.py / test_gpu_rollout.py) run at sizes the C oracle finishes in seconds.  Here the
packed-byte kernels (k_grouped_feats_x, k_rollout_x, info board from the feature kernel) are compared bit for bit with the
generic kernels they replaced (k_grouped_feats + info board from the step kernel, k_rollout) -- which are themselves
oracle-pinned and still serve the other board widths -- on 10^5 envs over whole games, so that rare situations (row clears in
odd places, stacks into the spawn rows, game overs, bag reshuffles) are hit many thousands of times.  The library reads its
kernel-selection switches (TG_GFEATS_V1, TG_INFO_IN_STEP, TG_ROLLOUT_V1) at every launch, so both variants run in one process.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        for k, v in self.kw.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


V1 = dict(TG_GFEATS_V1="1", TG_INFO_IN_STEP="1", TG_ROLLOUT_V1="1")
NEW = dict(TG_GFEATS_V1=None, TG_INFO_IN_STEP=None, TG_ROLLOUT_V1=None)


@pytest.mark.parametrize("W,H,n,T", [(10, 20, 120_000, 150), (20, 40, 20_000, 260), (10, 40, 30_000, 200)], ids=lambda v: str(v))
def test_grouped_features_packed_vs_generic_kernel(W, H, n, T):
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations

    def make():
        base = Tetris(width=W, height=H, gravity=False, queue_size=4, num_envs=n, autoreset_mode="next_step")
        return base, GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)], terminate_on_illegal_action=True)

    (ba, ea), (bb, eb) = make(), make()
    with _Env(**NEW):
        fa, ia = ea.reset(seed=123)
    with _Env(**V1):
        fb, ib = eb.reset(seed=123)
    assert torch.equal(fa, fb) and torch.equal(ia["action_mask"], ib["action_mask"]) and torch.equal(ia["board"], ib["board"])
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    lines = terms = 0
    for t in range(T):
        mask = ia["action_mask"].float()
        a = torch.multinomial(mask + 1e-9, 1, generator=g).squeeze(1).to(torch.int32)
        if t % 7 == 3:     # a sprinkle of arbitrary (often illegal) placements: terminate + `high` observation path
            r = torch.randint(0, 4 * W, (n,), device="cuda", generator=g, dtype=torch.int32)
            a = torch.where(torch.rand(n, device="cuda", generator=g) < 0.02, r, a)
        with _Env(**NEW):
            fa, ra, ta, _, ia = ea.step(a)
        with _Env(**V1):
            fb, rb, tb, _, ib = eb.step(a)
        assert torch.equal(fa, fb), t
        assert torch.equal(ia["action_mask"], ib["action_mask"]) and torch.equal(ia["board"], ib["board"]), t
        assert torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(ia["lines_cleared"], ib["lines_cleared"]), t
        lines += int(ia["lines_cleared"].sum())
        terms += int(ta.sum())
    sa, sb = ba.get_state(), bb.get_state()
    for k in sa:
        if torch.is_tensor(sa[k]):
            assert torch.equal(sa[k], sb[k]), k
    assert terms > n // 2, "the run should cover whole games"
    ba.close(); bb.close()


@pytest.mark.parametrize("W,H,n,K", [(10, 20, 100_000, 400), (20, 40, 8_000, 700), (10, 40, 20_000, 500)], ids=lambda v: str(v))
def test_rollout_packed_vs_generic_kernel(W, H, n, K):
    from tetris_gymnasium_b200.envs.tetris import Tetris

    res = []
    for flags in (NEW, V1):
        env = Tetris(width=W, height=H, gravity=False, queue_size=7, num_envs=n, autoreset_mode="next_step")
        env.reset(seed=77)
        with _Env(**flags):
            env.rollout((-51, 76, -36, -18), K // 2)
            last = env.rollout((-51, 76, -36, -18), K - K // 2, trace=True)
        st = env.get_state()
        res.append((st, last.clone(), env.episode_stats()))
        env.close()
    (sa, la, ea), (sb, lb, eb) = res
    for k in sa:
        if torch.is_tensor(sa[k]):
            assert torch.equal(sa[k], sb[k]), k
    assert torch.equal(la, lb)
    assert all(torch.equal(ea[k], eb[k]) for k in ea), (ea, eb)
    if H == 20:   # (statistics count finished episodes only; on the tall boards the policy does not lose within K steps)
        assert float(ea["sum_lines"]) > n, "the heuristic policy should clear many rows"


def test_rollout_weights_that_lose_quickly_packed_vs_generic_kernel():
    """A policy that stacks as high as possible: many game overs, placements ending in the spawn rows, resets."""
    from tetris_gymnasium_b200.envs.tetris import Tetris

    res = []
    for flags in (NEW, V1):
        env = Tetris(num_envs=60_000, gravity=False, queue_size=4, autoreset_mode="same_step")
        env.reset(seed=3)
        with _Env(**flags):
            env.rollout((9, -5, 7, 3), 300)
        res.append((env.get_state(), env.episode_stats()))
        env.close()
    (sa, ea), (sb, eb) = res
    for k in sa:
        if torch.is_tensor(sa[k]):
            assert torch.equal(sa[k], sb[k]), k
    assert all(torch.equal(ea[k], eb[k]) for k in ea), (ea, eb)
    assert float(ea["episodes"]) > 60_000

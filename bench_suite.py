#!/usr/bin/env python
"""bench_suite.py -- the other BASELINE.json configs (bench.py carries the headline config 2).

  C1 single env latency (reference config 1), C2 sweep over env counts, C3 grouped + features,
  C4 fused heuristic rollout, C5 wide board + RGB image.  One JSON object per line; `--out` also writes a
  markdown table.  CUDA-event timing, >= 3 warm-up iterations, inputs larger than L2 at the large sizes.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from tetris_gymnasium_b200.envs.tetris import Tetris  # noqa: E402
from tetris_gymnasium_b200.wrappers import CnnObservation, FeatureVectorObservation, GroupedActionsObservations, RgbObservation  # noqa: E402

PEAK = 6553.3
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def c2_sweep(out):
    for n in (4096, 16384, 65536, 262144, 1 << 20, 1 << 22):
        env = Tetris(num_envs=n, queue_size=7)
        env.reset(seed=42)
        K = 100 if n <= (1 << 20) else 30
        acts = torch.randint(0, 8, (K + 3, n), dtype=torch.int32, device="cuda")
        dt = timed(lambda i=0: env.step(acts[i]), K)
        lay = env.layout
        bps = lay.hot_stride * 2 + lay.board_stride * 1.15 + lay.rng_stride + 4 + 10 + 2 * lay.obs_board_bytes + 16 + 16 * 7
        out({"config": "C2 per-call step 10x20 q7, obs dict", "envs": n, "ms": dt * 1e3, "env_steps_per_s": n / dt,
             "GBps": bps * n / dt / 1e9, "frac_of_hbm_peak": bps * n / dt / 1e9 / PEAK})
        env.close()


def c2n_multi_step(out):
    """Config 2 at its small end through tg_step_n: K steps per native call, records resident in shared memory, obs dict +
    5-tuple of EVERY step written to [K][n] rollout storage.  HBM bytes per env-step: dict + 10 B outputs + 4 B action."""
    for n in (4096, 16384, 65536, 131072):
        env = Tetris(num_envs=n, queue_size=7)
        env.reset(seed=42)
        K = 64
        acts = torch.randint(0, 8, (4, K, n), dtype=torch.int32, device="cuda")
        dt = timed(lambda i=0: env.step_n(acts[i % 4]), 10) / K
        lay = env.layout
        bps = 2 * lay.obs_board_bytes + 16 + 16 * 7 + 10 + 4
        out({"config": f"C2n tg_step_n 10x20 q7, K={K} steps per call, obs dict of every step", "envs": n, "ms": dt * 1e3, "us_per_step": dt * 1e6,
             "env_steps_per_s": n / dt, "GBps": bps * n / dt / 1e9, "frac_of_hbm_peak": bps * n / dt / 1e9 / PEAK})
        env.close()


def c1_latency(out):
    env = Tetris(num_envs=1, queue_size=7)
    env.reset(seed=42)
    acts = torch.randint(0, 8, (1003, 1), dtype=torch.int32, device="cuda")
    dt = timed(lambda i=0: env.step(acts[i]), 1000)
    out({"config": "C1 single env, random actions (launch latency bound)", "envs": 1, "ms": dt * 1e3, "env_steps_per_s": 1 / dt})


def c3_grouped(out):
    for n in (4096, 65536, 1 << 20):
        base = Tetris(num_envs=n, gravity=False, queue_size=4)
        env = GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)])
        feats, info = env.reset(seed=42)
        K = 50

        def step(i=0):
            mask = env.legal_actions_mask.float()
            a = torch.multinomial(mask + 1e-9, 1).squeeze(1).to(torch.int32)     # uniformly random legal placement
            env.step(a)
        # time the env call alone: pre-sample actions from the current mask each iteration outside the events is not
        # possible (mask changes), so report both with and without the sampling kernel
        dt_all = timed(step, K)
        a = torch.multinomial(env.legal_actions_mask.float() + 1e-9, 1).squeeze(1).to(torch.int32)
        dt_env = timed(lambda i=0: env.step(a), K)   # same placement id every step (may become illegal -> handled)
        A, F = 40, 13
        bps = 2 * (32 + 144 + 16) + 32 + 144 * 1.0 + A * F + 2 * A + F + 10 + 4
        out({"config": "C3 grouped + features 10x20 (gravity off)", "envs": n, "ms": dt_env * 1e3, "env_steps_per_s": n / dt_env,
             "placements_per_s": n * A / dt_env, "with_action_sampling_env_steps_per_s": n / dt_all,
             "GBps": bps * n / dt_env / 1e9, "frac_of_hbm_peak": bps * n / dt_env / 1e9 / PEAK})
        base.close()
    for n, kw, name in ((65536, dict(queue_size=4), "10x20"), (262144, dict(queue_size=4), "10x20"),
                        (32768, dict(width=20, height=40, queue_size=5), "20x40")):
        base = Tetris(num_envs=n, gravity=False, **kw)
        env = GroupedActionsObservations(base)
        env.reset(seed=42)
        a = torch.zeros(n, dtype=torch.int32, device="cuda") + 17
        dt = timed(lambda i=0: env.step(a), 20)
        lay = base.layout
        A = lay.n_placements
        # placement images + legal mask written; state read + written by the placement step; obs dict of the step
        bps = A * lay.obs_board_bytes + A + 2 * (lay.hot_stride + lay.board_stride) + lay.board_stride + lay.rng_stride + 2 * lay.obs_board_bytes + 16 + lay.obs_queue_bytes + 14
        out({"config": f"C3b grouped boards {name} (no wrappers)", "envs": n, "ms": dt * 1e3, "env_steps_per_s": n / dt,
             "placements_per_s": n * A / dt, "GBps": bps * n / dt / 1e9, "frac_of_hbm_peak": bps * n / dt / 1e9 / PEAK})
        base.close()


def c4_rollout(out):
    for n in (1 << 18, 1 << 21):
        env = Tetris(num_envs=n, gravity=False, queue_size=7)
        env.reset(seed=42)
        env.rollout((-51, 76, -36, -18), 16)
        K = 256
        dt = timed(lambda i=0: env.rollout((-51, 76, -36, -18), K), 2, warm=1)
        st = {k: float(v) for k, v in env.episode_stats().items()}
        out({"config": "C4 fused heuristic rollout 10x20 q7, K=256 per launch", "envs": n, "ms_per_launch": dt * 1e3,
             "env_steps_per_s": n * K / dt, "placements_per_s": n * K * 40 / dt, "episode_stats": st})
        env.close()


def c5_wide_rgb(out):
    for n in (65536, 1 << 18):
        base = Tetris(num_envs=n, width=20, height=40, queue_size=5)
        env = RgbObservation(base)
        env.reset(seed=42)
        K = 50
        acts = torch.randint(0, 8, (K + 3, n), dtype=torch.int32, device="cuda")
        dt = timed(lambda i=0: env.step(acts[i]), K)
        lay = base.layout
        img = lay.height_padded * lay.rgb_width * 3
        bps = 2 * (lay.hot_stride + lay.board_stride * 1.0 + lay.rng_stride) + lay.hot_stride + 0.1 * lay.board_stride + img + 14
        out({"config": "C5 wide 20x40 q5 + RGB image obs", "envs": n, "image_bytes": img, "ms": dt * 1e3, "env_steps_per_s": n / dt,
             "GBps": bps * n / dt / 1e9, "frac_of_hbm_peak": bps * n / dt / 1e9 / PEAK})
        base.close()
    for n in (65536, 1 << 18):
        base = Tetris(num_envs=n, width=20, height=40, queue_size=5)
        env = CnnObservation(base, shape=(84, 84), stack_size=4, window=28, clip_reward=True)
        env.reset(seed=42)
        K = 50
        acts = torch.randint(0, 8, (K + 3, n), dtype=torch.int32, device="cuda")
        dt = timed(lambda i=0: env.step(acts[i]), K)
        lay = base.layout
        # state read+written by the step, state read by the adapter, one 84x84 frame written (+ 3/28 of a frame for the window slide)
        bps = 2 * (lay.hot_stride + lay.board_stride * 1.0 + lay.rng_stride) + lay.hot_stride + 0.1 * lay.board_stride + 84 * 84 * (1 + 2 * 3 / 28) + 14
        out({"config": "C5c wide 20x40 q5 + fused CNN obs (resize 84x84, grey, stack 4)", "envs": n, "frame_bytes": 84 * 84, "ms": dt * 1e3,
             "env_steps_per_s": n / dt, "GBps": bps * n / dt / 1e9, "frac_of_hbm_peak": bps * n / dt / 1e9 / PEAK})
        base.close()
    base = Tetris(num_envs=1 << 18, width=20, height=40, queue_size=5)
    base.reset(seed=42)
    acts = torch.randint(0, 8, (53, 1 << 18), dtype=torch.int32, device="cuda")
    dt = timed(lambda i=0: base.step(acts[i]), 50)
    lay = base.layout
    bps = 2 * lay.hot_stride + lay.board_stride * 1.1 + lay.rng_stride + 14 + 2 * lay.obs_board_bytes + 16 + 80
    out({"config": "C5b wide 20x40 q5, obs dict", "envs": 1 << 18, "ms": dt * 1e3, "env_steps_per_s": (1 << 18) / dt,
         "GBps": bps * (1 << 18) / dt / 1e9, "frac_of_hbm_peak": bps * (1 << 18) / dt / 1e9 / PEAK})


def c6_functional(out):
    """Functional facade (envs/tetris_fn.py batched_step): explicit State in / out every call (board i8[24,18] + scalars)."""
    from tetris_gymnasium_b200.envs import tetris_fn as F
    from tetris_gymnasium_b200.functional.core import EnvConfig
    from tetris_gymnasium_b200.functional.tetrominoes import TETROMINOES

    cfg = EnvConfig(width=10, height=20, padding=4, queue_size=7)
    for n in (65536, 1 << 20):
        keys = torch.stack([torch.arange(n, device="cuda"), torch.full((n,), 42, device="cuda")], dim=1)
        keys, state, obs = F.batched_reset(TETROMINOES, keys, config=cfg)
        acts = torch.randint(0, 7, (24, n), dtype=torch.int32, device="cuda")
        box = {"s": state}

        def step(i=0):
            box["s"], _, _, _, _ = F.batched_step(TETROMINOES, box["s"], acts[i % 24], config=cfg)
        dt = timed(step, 20)
        bps = 2 * 432 + 200 + 2 * 4 * 16 + 4 + 9
        out({"config": "C6 functional facade batched_step 10x20 (State in/out every call, incl. the host-side State packing)", "envs": n,
             "ms": dt * 1e3, "env_steps_per_s": n / dt, "GBps": bps * n / dt / 1e9, "frac_of_hbm_peak": bps * n / dt / 1e9 / PEAK})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    rows = []

    def out(d):
        d["gpu"] = torch.cuda.get_device_name(0)
        rows.append(d)
        print(json.dumps(d), flush=True)

    for name, fn in (("c1", c1_latency), ("c2", c2_sweep), ("c2n", c2n_multi_step), ("c3", c3_grouped), ("c4", c4_rollout), ("c5", c5_wide_rgb), ("c6", c6_functional)):
        if not args.only or name in args.only.split(","):
            fn(out)
    if args.out:
        with open(args.out + ".json", "w") as f:
            json.dump(rows, f, indent=1)
        with open(args.out + ".md", "w") as f:
            f.write("| config | envs | ms | env-steps/s | placements/s | GB/s | of HBM peak |\n|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| {r['config']} | {r['envs']} | {r.get('ms', r.get('ms_per_launch', 0)):.3f} | {r['env_steps_per_s']:.4g} | "
                        f"{r.get('placements_per_s', 0):.4g} | {r.get('GBps', 0):.0f} | {100 * r.get('frac_of_hbm_peak', 0):.1f}% |\n")


if __name__ == "__main__":
    main()

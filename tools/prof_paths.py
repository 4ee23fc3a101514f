"""Drive one path of the library for a few iterations so that ncu can capture its kernel.

    ncu --set full --clock-control none --import-source on -k regex:k_rgb -s 4 -c 1 -o gpurun_out/prof_rgb \
        python tools/prof_paths.py rgb --envs 262144

paths: step | cnn (wide board, fused 84x84 grey frame stack) | rgb (wide board + image) | rgb_d (default board + image) | boards | boards_x | feats | rollout | fn (functional facade batched_step)"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from tetris_gymnasium_b200.envs.tetris import Tetris  # noqa: E402
from tetris_gymnasium_b200.wrappers import CnnObservation, FeatureVectorObservation, GroupedActionsObservations, RgbObservation  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("path")
ap.add_argument("--envs", type=int, default=1 << 18)
ap.add_argument("--iters", type=int, default=12)
args = ap.parse_args()
n = args.envs
if args.path == "step":
    env = Tetris(num_envs=n, queue_size=7)
    env.reset(seed=42)
    acts = torch.randint(0, 8, (args.iters, n), dtype=torch.int32, device="cuda")
    for i in range(args.iters):
        env.step(acts[i])
elif args.path in ("rgb", "rgb_d"):
    base = Tetris(num_envs=n, width=20, height=40, queue_size=5) if args.path == "rgb" else Tetris(num_envs=n, queue_size=4)
    env = RgbObservation(base)
    env.reset(seed=42)
    acts = torch.randint(0, 8, (args.iters, n), dtype=torch.int32, device="cuda")
    for i in range(args.iters):
        env.step(acts[i])
elif args.path in ("boards", "boards_x", "feats"):
    base = Tetris(num_envs=n, width=20, height=40, queue_size=5, gravity=False) if args.path == "boards_x" else Tetris(num_envs=n, gravity=False, queue_size=4)
    wr = [FeatureVectorObservation(base)] if args.path == "feats" else None
    env = GroupedActionsObservations(base, observation_wrappers=wr)
    env.reset(seed=42)
    for i in range(args.iters):
        a = torch.multinomial(env.legal_actions_mask.float() + 1e-9, 1).squeeze(1).to(torch.int32)
        env.step(a)
elif args.path == "cnn":
    base = Tetris(num_envs=n, width=20, height=40, queue_size=5)
    env = CnnObservation(base)
    env.reset(seed=42)
    acts = torch.randint(0, 8, (args.iters, n), dtype=torch.int32, device="cuda")
    for i in range(args.iters):
        env.step(acts[i])
elif args.path == "rollout":
    env = Tetris(num_envs=n, gravity=False, queue_size=7)
    env.reset(seed=42)
    for i in range(3):
        env.rollout((-51, 76, -36, -18), 64)
elif args.path == "fn":
    from tetris_gymnasium_b200.envs import tetris_fn as F
    from tetris_gymnasium_b200.functional.core import EnvConfig
    from tetris_gymnasium_b200.functional.tetrominoes import TETROMINOES
    cfg = EnvConfig(width=10, height=20, padding=4, queue_size=7)
    keys = torch.stack([torch.arange(n, device="cuda"), torch.full((n,), 42, device="cuda")], dim=1)
    keys, state, obs = F.batched_reset(TETROMINOES, keys, config=cfg)
    acts = torch.randint(0, 7, (args.iters, n), dtype=torch.int32, device="cuda")
    for i in range(args.iters):
        state, _, _, _, _ = F.batched_step(TETROMINOES, state, acts[i], config=cfg)
torch.cuda.synchronize()

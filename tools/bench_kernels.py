"""Kernel-only timings (CUDA events) of the wrapper kernels, next to a write-only memset of the same size.

    python tools/bench_kernels.py boards|boards_x|rgb|rgb_d|feats [--envs N]"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from tetris_gymnasium_b200.envs.tetris import Tetris  # noqa: E402
from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations, RgbObservation  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


ap = argparse.ArgumentParser()
ap.add_argument("path")
ap.add_argument("--envs", type=int, default=1 << 17)
ap.add_argument("--play", type=int, default=12, help="random steps before timing (non-trivial boards)")
args = ap.parse_args()
n = args.envs
wide = args.path.endswith("_x") or args.path == "rgb"
kw = dict(width=20, height=40, queue_size=5) if wide else dict(queue_size=4)
if args.path.startswith("boards") or args.path == "feats":
    base = Tetris(num_envs=n, gravity=False, **kw)
    env = GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)] if args.path == "feats" else None)
    env.reset(seed=42)
    for i in range(args.play):
        a = torch.multinomial(env.legal_actions_mask.float() + 1e-9, 1).squeeze(1).to(torch.int32)
        env.step(a)
    out = env.observation()
    dt = timed(lambda: env.observation())
else:
    base = Tetris(num_envs=n, **kw)
    env = RgbObservation(base)
    env.reset(seed=42)
    acts = torch.randint(0, 8, (args.play, n), dtype=torch.int32, device="cuda")
    for i in range(args.play):
        env.step(acts[i])
    out = env.observation()
    dt = timed(lambda: env.observation())
nbytes = out.numel() * out.element_size()
buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
dt_set = timed(lambda: buf.fill_(1))
src = torch.empty(nbytes // 2, dtype=torch.uint8, device="cuda")
dt_cp = timed(lambda: buf[: nbytes // 2].copy_(src))
print(f"{args.path} envs={n} out={nbytes/1e9:.3f} GB  kernel {dt*1e3:.3f} ms = {nbytes/dt/1e9:.0f} GB/s written | "
      f"memset same bytes {dt_set*1e3:.3f} ms = {nbytes/dt_set/1e9:.0f} GB/s | copy {nbytes/dt_cp/1e9:.0f} GB/s (r+w)")

"""Compact host-buffer step (tg_step_host, TG_HOST_COMPACT) at the bench size for several chunk sizes / ring depths:
python tools/time_host.py [envs]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tetris_gymnasium_b200.envs.tetris import Tetris  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
env = Tetris(num_envs=n, queue_size=7)
env.reset(seed=42)
bufs = env.alloc_host_buffers(pinned=True)
rng = np.random.default_rng(0)
acts = [torch.from_numpy(rng.integers(0, 8, size=n).astype(np.int32)).pin_memory().numpy() for _ in range(6)]
for chunk, ring in ((131072, 4), (65536, 4), (32768, 4), (16384, 4), (16384, 8), (8192, 8), (32768, 8), (131072, 4)):
    os.environ["TG_HOST_CHUNK"] = str(chunk); os.environ["TG_HOST_RING"] = str(ring)
    for t in range(2):
        env.step_host(acts[t], bufs, mode="compact")
    K = 8
    t0 = time.perf_counter()
    w = x = 0.0
    for t in range(K):
        env.step_host(acts[t % 6], bufs, mode="compact")
        st = env.host_stats(); w += st["wait_s"]; x += st["expand_s"]
    dt = (time.perf_counter() - t0) / K
    print(f"chunk {chunk:7d} ring {ring}: {dt * 1e3:.2f} ms/step = {n / dt / 1e6:.1f} M env-steps/s (waiting {w / K * 1e3:.2f} ms, expanding {x / K * 1e3:.2f} ms)", flush=True)

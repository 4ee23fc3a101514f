"""Per-call time of the functional facade's batched_step (CUDA events and host wall clock):  python tools/time_fn.py"""
import sys
import time

import torch

sys.path.insert(0, ".")
from tetris_gymnasium_b200.envs import tetris_fn as F  # noqa: E402
from tetris_gymnasium_b200.functional.core import EnvConfig  # noqa: E402
from tetris_gymnasium_b200.functional.tetrominoes import TETROMINOES  # noqa: E402

cfg = EnvConfig(width=10, height=20, padding=4, queue_size=7)
for n in (4096, 65536, 262144, 1 << 20):
    keys = torch.stack([torch.arange(n, device="cuda"), torch.full((n,), 42, device="cuda")], dim=1)
    keys, state, obs = F.batched_reset(TETROMINOES, keys, config=cfg)
    acts = torch.randint(0, 7, (24, n), dtype=torch.int32, device="cuda")
    for i in range(30):
        state, _, _, _, _ = F.batched_step(TETROMINOES, state, acts[i % 24], config=cfg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 50
    t0 = time.perf_counter()
    e0.record()
    for i in range(K):
        state, _, _, _, _ = F.batched_step(TETROMINOES, state, acts[i % 24], config=cfg)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"envs {n}: device {e0.elapsed_time(e1) / K * 1e3:.1f} us/call, host enqueue {(t1 - t0) / K * 1e6:.1f} us/call")

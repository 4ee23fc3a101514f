"""In-stream timing of the grouped path's kernels (CUDA events, warm, back to back):  python tools/time_grouped.py [envs]
observe = k_grouped_feats_x alone (tg_grouped_observe on a steady-state population), step = tg_grouped_step (placement + enumeration)."""
import sys

import torch

sys.path.insert(0, ".")
from tetris_gymnasium_b200.envs.tetris import Tetris  # noqa: E402
from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
base = Tetris(num_envs=n, gravity=False, queue_size=4)
env = GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)])
env.reset(seed=42)


def sample():
    return torch.multinomial(env.legal_actions_mask.float() + 1e-9, 1).squeeze(1).to(torch.int32)


for _ in range(40):
    env.step(sample())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for _ in range(K):
    env.observation()
e1.record()
torch.cuda.synchronize()
t_obs = e0.elapsed_time(e1) / K * 1e3
acts = [sample() for _ in range(1)]
ts = []
for _ in range(K):
    a = sample()
    e0.record()
    env.step(a)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print(f"envs {n}: observe alone {t_obs:.1f} us; step (placement + enumeration) median {ts[len(ts) // 2]:.1f} us, min {ts[0]:.1f} us "
      f"-> {n * 40 / ts[len(ts) // 2] / 1e3:.1f} G placements/s")

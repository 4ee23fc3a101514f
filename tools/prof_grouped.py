import sys, torch
sys.path.insert(0, ".")
from tetris_gymnasium_b200.envs.tetris import Tetris
from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations
n = 1 << 20
base = Tetris(num_envs=n, gravity=False, queue_size=4)
env = GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)])
env.reset(seed=42)
for i in range(12):
    a = torch.multinomial(env.legal_actions_mask.float() + 1e-9, 1).squeeze(1).to(torch.int32)
    env.step(a)
torch.cuda.synchronize()

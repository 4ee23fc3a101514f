import sys, torch
sys.path.insert(0, ".")
from tetris_gymnasium_b200.envs.tetris import Tetris
from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations
kind, n, T = sys.argv[1], int(sys.argv[2]), 40
g = torch.Generator(device="cuda").manual_seed(5)
A = 40 if kind == "grouped" else 8
acts = torch.randint(0, A, (2, T, n), dtype=torch.int32, device="cuda", generator=g)
def make(seed):
    if kind == "grouped":
        base = Tetris(num_envs=n, gravity=False, queue_size=4)
        env = GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)])
    else:
        base = Tetris(num_envs=n, queue_size=7); env = base
    env.reset(seed=seed)
    return base, env
def run(concurrent):
    envs = [make(11), make(12)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()] if concurrent else [torch.cuda.current_stream()] * 2
    torch.cuda.synchronize()
    for t in range(T):
        for i, (base, env) in enumerate(envs):
            with torch.cuda.stream(streams[i]):
                env.step(acts[i, t])
    torch.cuda.synchronize()
    return [(b._hot.clone(), b._brd.clone()) for b, e in envs]
ref, con = run(False), run(True)
ok = all(torch.equal(a, b) for i in range(2) for a, b in zip(ref[i], con[i]))
print(kind, n, "two streams == serial:", ok)

"""GPU: two env handles stepped concurrently on two CUDA streams give exactly what they give one after the other (every
entry point is asynchronous on the caller's stream and keeps no state outside its handle and the caller's arrays)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(n, seed):
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations

    base = Tetris(num_envs=n, gravity=False, queue_size=4)
    env = GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)])
    env.reset(seed=seed)
    return base, env


@pytest.mark.parametrize("n", [4096 + 7, 262144])
def test_two_grouped_envs_on_two_streams(n):
    T = 40
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = torch.randint(0, 40, (2, T, n), dtype=torch.int32, device="cuda", generator=g)   # arbitrary placements, many illegal

    def run(concurrent):
        envs = [_make(n, 11), _make(n, 12)]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()] if concurrent else [torch.cuda.current_stream()] * 2
        torch.cuda.synchronize()
        for t in range(T):
            for i, (base, env) in enumerate(envs):
                with torch.cuda.stream(streams[i]):
                    env.step(acts[i, t])
        torch.cuda.synchronize()
        out = [(b._hot.clone(), b._brd.clone(), e._feats.clone(), e._legal.clone(), b._reward.clone()) for b, e in envs]
        for b, _ in envs:
            b.close()
        return out

    ref, con = run(False), run(True)
    for i in range(2):
        for a, b, name in zip(ref[i], con[i], ("hot", "board", "features", "legal mask", "reward")):
            assert torch.equal(a, b), f"env {i}: {name} differs between serial and two-stream execution"


def test_base_env_on_side_stream_matches_default_stream():
    from tetris_gymnasium_b200.envs.tetris import Tetris

    n, T = 65536 + 3, 30
    g = torch.Generator(device="cuda").manual_seed(6)
    acts = torch.randint(0, 8, (T, n), dtype=torch.int32, device="cuda", generator=g)

    def run(stream):
        env = Tetris(num_envs=n, queue_size=7)
        with torch.cuda.stream(stream):
            env.reset(seed=3)
            for t in range(T):
                obs, r, term, trunc, info = env.step(acts[t])
        stream.synchronize()
        out = {k: v.clone() for k, v in obs.items()}
        out["hot"] = env._hot.clone()
        env.close()
        return out

    torch.cuda.synchronize()
    a, b = run(torch.cuda.current_stream()), run(torch.cuda.Stream())
    for k in a:
        assert torch.equal(a[k], b[k]), k

"""GPU: the drop-in surface around the kernels -- wrapper-level options of the reference (terminate_on_illegal_action on the
wrapper alone, float64 legal mask / float observations), the vector-env info format (`_key` masks), the per-env invalid-action
flags, RecordEpisodeStatistics under every autoreset mode, generic observation wrappers, the device guard."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _grouped(n, seqs, terminate, W=10, H=20, Q=4, **kw):
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations

    base = Tetris(width=W, height=H, gravity=False, queue_size=Q, num_envs=n, randomizer_mode="sequence", piece_sequences=seqs,
                  autoreset_mode="disabled")        # NOTE: terminate_on_illegal_action only on the wrapper (reference :44-49)
    return base, GroupedActionsObservations(base, observation_wrappers=[FeatureVectorObservation(base)], terminate_on_illegal_action=terminate, **kw)


@pytest.mark.parametrize("terminate", [True, False])
def test_terminate_on_illegal_is_a_wrapper_option(terminate):
    from oracle.tetris_oracle import OracleEnv
    from gpu_util import np_

    n = 64
    rng = np.random.default_rng(4)
    seqs = rng.integers(0, 7, size=(n, 64)).astype(np.uint8)
    base, env = _grouped(n, seqs, terminate)
    orc = [OracleEnv(gravity=False) for _ in range(n)]
    for i, o in enumerate(orc):
        o.set_sequence(seqs[i]); o.reset(); o.grouped_observe()
    g, info = env.reset()
    a = np.zeros(n, np.int64)                        # action 0 is illegal for most pieces (column 0, rotation 0)
    legal0 = np_(info["action_mask"])[:, 0].astype(bool)
    assert (~legal0).any()
    g, r, term, trunc, info = env.step(torch.from_numpy(a))
    for i, o in enumerate(orc):
        code, rr, tt, ll = o.grouped_step(0, terminate_on_illegal=terminate)
        assert np.float32(rr) == np_(r)[i] and bool(tt) == bool(np_(term)[i]), i
    if terminate:
        assert np_(term)[~legal0].all() and (np_(g)[~legal0] == 200).all()      # observation_space.high = H * W
    else:
        assert not np_(term)[~legal0].any() and (np_(r)[~legal0] == np.float32(-0.1)).all()


def test_mask_and_obs_dtypes_and_saturated_fill_on_the_wide_board():
    from gpu_util import np_

    n = 32
    seqs = np.zeros((n, 16), np.uint8)               # I pieces: column 0 / rotation 0 is illegal
    base, env = _grouped(n, seqs, True, W=20, H=40, Q=5)
    g, info = env.reset()
    assert g.dtype == torch.uint8 and info["action_mask"].dtype == torch.uint8
    g, r, term, trunc, info = env.step(torch.zeros(n, dtype=torch.int64))
    assert np_(term).all() and (np_(g) == 255).all()  # H * W = 800 does not fit uint8: saturated, not wrapped (800 & 255 = 32)
    base2, env2 = _grouped(n, seqs, True, W=20, H=40, Q=5, mask_dtype=torch.float64, obs_dtype=torch.float32)
    g2, info2 = env2.reset()
    assert g2.dtype == torch.float32 and info2["action_mask"].dtype == torch.float64 and env2.legal_actions_mask.dtype == torch.float64
    assert np.array_equal(np_(info2["action_mask"]), np_(info["action_mask"]).astype(np.float64) * 0 + np_(env2.legal_actions_mask))
    g2, r2, term2, _, info2 = env2.step(torch.zeros(n, dtype=torch.int64))
    assert np_(term2).all() and (np_(g2) == 800.0).all()     # a float observation carries the reference's value


def test_vector_info_format_and_invalid_action_flags():
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from gpu_util import np_

    n = 100
    env = Tetris(num_envs=n, report_invalid_actions=True)
    obs, info = env.reset(seed=1)
    assert "lines_cleared" in info and info["_lines_cleared"].dtype == torch.bool and bool(info["_lines_cleared"].all())
    a = torch.randint(0, 8, (n,), dtype=torch.int32, device="cuda")
    a[3], a[17], a[50] = 8, -1, 12345
    twin = Tetris(num_envs=n)
    twin.reset(seed=1)
    b = a.clone()
    b[3] = b[17] = b[50] = 7                         # an out-of-range action behaves like the unmatched elif chain: no move
    obs, r, term, trunc, info = env.step(a)
    o2, r2, _, _, _ = twin.step(b)
    flags = np_(info["invalid_action"])
    assert flags.sum() == 3 and flags[[3, 17, 50]].all() and bool(info["_invalid_action"].all())
    for k in obs:
        assert torch.equal(obs[k], o2[k]), k


@pytest.mark.parametrize("mode", ["next_step", "same_step", "disabled"])
def test_record_episode_statistics_under_every_autoreset_mode(mode):
    from oracle.tetris_oracle import OracleEnv
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import RecordEpisodeStatistics
    from gpu_util import np_

    n = 48
    rng = np.random.default_rng(8)
    seqs = rng.integers(0, 7, size=(n, 64)).astype(np.uint8)
    env = RecordEpisodeStatistics(Tetris(num_envs=n, randomizer_mode="sequence", piece_sequences=seqs, autoreset_mode=mode))
    env.reset()
    ret, length, pending, done = np.zeros(n), np.zeros(n, int), np.zeros(n, bool), np.zeros(n, bool)
    seen = 0
    for t in range(400):
        a = rng.choice([5, 5, 5, 0, 1, 6], size=n)
        obs, r, term, trunc, info = env.step(torch.from_numpy(a))
        r, term = np_(r).astype(np.float64), np_(term)
        for i in range(n):
            if mode == "disabled" and done[i]:
                continue
            if mode == "next_step" and pending[i]:
                pending[i] = False
                ret[i], length[i] = 0.0, 0
                continue
            ret[i] += r[i]; length[i] += 1
            if term[i]:
                seen += 1
                assert bool(np_(info["_episode"])[i])
                assert abs(float(np_(info["episode"]["r"])[i]) - ret[i]) < 1e-3 and int(np_(info["episode"]["l"])[i]) == length[i], (mode, t, i)
                if mode == "next_step":
                    pending[i] = True
                elif mode == "same_step":
                    ret[i], length[i] = 0.0, 0
                else:
                    done[i] = True
    assert seen > n // 2


def test_generic_observation_wrappers_are_applied_per_placement():
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.wrappers import GroupedActionsObservations
    from gpu_util import np_

    class CellCount:                                  # any object with observation(dict) -> tensor, like a gym.ObservationWrapper
        def observation(self, obs):
            b = obs["board"][:, :20, 4:14]
            return torch.stack([(b > 0).sum(dim=(1, 2)), obs["queue"].sum(dim=(1, 2)).to(torch.int64)], dim=1)

    n = 16
    base = Tetris(num_envs=n, gravity=False)
    env = GroupedActionsObservations(base, observation_wrappers=[CellCount(), CellCount.__new__(CellCount)][:1])
    g, info = env.reset(seed=5)
    plain = GroupedActionsObservations(Tetris(num_envs=n, gravity=False))
    boards, _ = plain.reset(seed=5)
    assert g.shape == (n, 40, 2)
    want = (np_(boards)[:, :, :20, 4:14] > 0).sum(axis=(2, 3))
    assert np.array_equal(np_(g)[:, :, 0], want)


def test_entry_points_leave_the_current_device_alone():
    from tetris_gymnasium_b200.envs.tetris import Tetris

    cur = torch.cuda.current_device()
    env = Tetris(num_envs=8, randomizer_mode="numpy")
    env.reset(seed=3)
    env.get_state(); env.step(torch.zeros(8, dtype=torch.int32, device="cuda")); env.step_host(np.zeros(8, np.int32))
    env.close()
    assert torch.cuda.current_device() == cur

#!/usr/bin/env python
"""bench.py -- headline benchmark: batched per-call Tetris env step on B200 (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our CUDA arm
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port of the reference on host cores

A "step" = one `env.step(actions)` over ENVS_PER_GPU envs of the default 10x20 board, uniformly random
actions over all 8 ids, observation dict (board + active mask + holder + queue) written every step,
NEXT_STEP autoreset, device-native Philox 7-bag.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, QUEUE = 10, 20, 7          # BASELINE.json: "default 10x20 board, padding 4, queue_size 7"
ENVS_PER_GPU = 1 << 22                    # 4,194,304 envs/GPU (top of BASELINE's "4K to 4M envs"): obs dict 4.2 GB per step (>> 126 MB L2)
METRIC, UNIT = "env-steps/s (batched per-call step, 10x20, obs dict every step)", "env-steps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bytes_per_step(layout, queue, commit_frac):
    """Algorithmic (compulsory) HBM bytes per env-step of OUR layout ('Roofline')."""
    ob = layout.obs_board_bytes
    obs = 2 * ob + 16 + 16 * queue
    # rng record: read every step; written back when a bag is reshuffled (one commit in seven)
    read = layout.hot_stride + layout.board_stride + layout.rng_stride + 4
    write = layout.hot_stride + obs + 10 + commit_frac * layout.board_stride + commit_frac / 7.0 * layout.rng_stride
    return read + write, obs


def host_cores():
    """Host threads the CPU arm uses: every core this process may run on (torchrun exports OMP_NUM_THREADS=1 to its
    workers, so the OpenMP default is not trusted; the thread count is passed to the oracle explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(seconds_target=12.0, n_envs=8192):
    """The oracle port (oracle/tetris_oracle.c, OpenMP over host cores) on a bounded sample of the same workload."""
    from oracle.tetris_oracle import OracleVec, lib

    cores = host_cores()
    vec = OracleVec(n_envs, width=WIDTH, height=HEIGHT, gravity=True, queue_size=QUEUE)
    for i, e in enumerate(vec.envs):
        e.seed_numpy(1 + i)
        e.reset()
    rng = np.random.default_rng(42)
    acts = rng.integers(0, 8, size=(64, n_envs)).astype(np.int32)
    for t in range(3):
        vec.step(acts[t], nthreads=cores)
    t0 = time.perf_counter()
    steps = 0
    while True:
        vec.step(acts[steps % 64], nthreads=cores)
        steps += 1
        if time.perf_counter() - t0 > seconds_target:
            break
    dt = time.perf_counter() - t0
    return {"value": n_envs * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_envs} envs x {steps} vector steps ({dt:.1f} s), oracle/tetris_oracle.c via OpenMP, NEXT_STEP autoreset, numpy-exact 7-bag"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.tetris_oracle import OracleVec, lib

    cores = host_cores()
    n_envs = 16384
    vec = OracleVec(n_envs, width=WIDTH, height=HEIGHT, gravity=True, queue_size=QUEUE)
    for i, e in enumerate(vec.envs):
        e.seed_numpy(1 + i)
        e.reset()
    rng = np.random.default_rng(42)
    acts = rng.integers(0, 8, size=(args.warmup + args.steps, n_envs)).astype(np.int32)
    for t in range(args.warmup):
        vec.step(acts[t], nthreads=cores)
    t0 = time.perf_counter()
    for t in range(args.steps):
        vec.step(acts[args.warmup + t], nthreads=cores)
    dt = time.perf_counter() - t0
    v = n_envs * args.steps / dt
    sample = f"{n_envs} envs per step (bounded sample of the {ENVS_PER_GPU}-env workload), oracle port of the reference NumPy env, OpenMP x{cores}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"per-call step {WIDTH}x{HEIGHT} queue {QUEUE}, random actions, obs dict every step", "envs_per_step": n_envs},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the grouped / rollout side measurements")
    ap.add_argument("--no-probe", action="store_true", help="skip the commit-fraction probe (torch kernels) e.g. under ncu")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from tetris_gymnasium_b200.envs.tetris import Tetris

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from tetris_gymnasium_b200.sharding import allreduce_episode_stats, shard_range

    n, K, Wm = args.envs, args.steps, args.warmup
    start, stop = shard_range(world * n, rank, world)      # weak scaling: n envs per GPU, global ids keyed by shard
    env = Tetris(width=WIDTH, height=HEIGHT, gravity=True, queue_size=QUEUE, num_envs=stop - start, device=dev,
                 autoreset_mode="next_step", randomizer_mode="philox", env_id_offset=start)
    env.reset(seed=42)
    g = torch.Generator(device=dev)
    g.manual_seed(42 + rank)
    acts = torch.randint(0, 8, (Wm + K, n), dtype=torch.int32, device=dev, generator=g)
    for t in range(Wm):
        env.step(acts[t])
    env.episode_stats(reset=True)

    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0.record()
    for t in range(K):
        env.step(acts[Wm + t])
    allreduce_episode_stats(env._stats)   # the only collective: episode statistics (4 doubles) over NCCL
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms[0])
    value = world * n * K / (ms_max * 1e-3)

    # commit fraction (share of env-steps that write their board record back), probed AFTER the timed region
    commit_frac = 0.0
    if not args.no_probe:
        m = min(n, 65536)                      # a 64 K-env sample keeps the probe's torch kernels negligible
        prev_q = env._o_queue[:m].clone()
        changed = 0.0
        for t in range(8):
            env.step(acts[t % (Wm + K)])
            changed += float((env._o_queue[:m] != prev_q).flatten(1).any(1).float().mean())
            prev_q.copy_(env._o_queue[:m])
        commit_frac = changed / 8
    else:
        commit_frac = 0.148

    # end-to-end through host buffers (tg_step_host): pinned actions H2D, obs dict + 5-tuple D2H every step
    e2e = None
    if not args.no_e2e:
        bufs = env.alloc_host_buffers(pinned=True)
        Ke = min(K, 20)            # every e2e step moves the whole observation dict over PCIe
        h_acts = torch.empty((Ke + 2, n), dtype=torch.int32, pin_memory=True)
        h_acts.copy_(acts[:Ke + 2])
        h_np = h_acts.numpy()
        for t in range(2):
            env.step_host(h_np[t], bufs)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(Ke):
            env.step_host(h_np[2 + t], bufs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tdt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
        d2h = sum(int(np.prod(v.shape)) * v.dtype.itemsize for v in bufs.values())
        # context: the same loop when the policy lives on the GPU -- actions H2D, step, only the 5-tuple scalars D2H
        # (the observation dict stays in HBM, which is how the device API `env.step(cuda_actions)` is used)
        h_rew = torch.empty(n, dtype=torch.float32, pin_memory=True)
        h_term = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        d_act = torch.empty(n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for t in range(Ke):
            d_act.copy_(h_acts[2 + t], non_blocking=True)
            _, rew, term, _, _ = env.step(d_act)
            h_rew.copy_(rew, non_blocking=True)
            h_term.copy_(term.view(torch.uint8), non_blocking=True)
            torch.cuda.synchronize()
        dt_s = time.perf_counter() - t1
        tds = torch.tensor([dt_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tds, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n * Ke / float(tdt[0]), "unit": UNIT, "steps": Ke, "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": d2h,
               "pcie_gb_per_s_per_gpu": (4 * n + d2h) * Ke / float(tdt[0]) / 1e9,   # the host link, not the GPU, bounds this number
               "note": "tg_step_host: actions from pinned host memory, full observation dict + reward/terminated/truncated/lines read back to pinned host memory every step",
               "obs_on_device": {"value": world * n * Ke / float(tds[0]), "unit": UNIT, "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 5 * n,
                                 "note": "same loop with the observation dict left in HBM (GPU-resident policy): actions H2D, step, reward + terminated D2H, synchronised every step"}}

    # ---- the other half of BASELINE's metric: grouped placements/s (config 3) and the fused rollout (config 4), short runs ----
    extra = None
    layout = env.layout
    if not args.no_extra:
        from tetris_gymnasium_b200.wrappers import FeatureVectorObservation, GroupedActionsObservations
        del env, acts
        torch.cuda.empty_cache()
        ng = min(n, 1 << 20)
        gbase = Tetris(width=WIDTH, height=HEIGHT, gravity=False, queue_size=4, num_envs=ng, device=dev, env_id_offset=rank * ng)
        genv = GroupedActionsObservations(gbase, observation_wrappers=[FeatureVectorObservation(gbase)])
        genv.reset(seed=42)
        Kg, tg_ms = 24, 0.0
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for t in range(Kg + 4):
            a = torch.multinomial(genv.legal_actions_mask.float() + 1e-9, 1).squeeze(1).to(torch.int32)   # random legal placement (untimed)
            g0.record()
            genv.step(a)
            g1.record()
            torch.cuda.synchronize()
            if t >= 4:
                tg_ms += g0.elapsed_time(g1)
        Kr = 128
        gbase.rollout((-51, 76, -36, -18), 16)
        g0.record()
        gbase.rollout((-51, 76, -36, -18), Kr)
        g1.record()
        torch.cuda.synchronize()
        tr_ms = g0.elapsed_time(g1)
        tt = torch.tensor([tg_ms, tr_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        A = gbase.layout.n_placements
        extra = {"grouped": {"placements_per_s": world * ng * A * Kg / (float(tt[0]) * 1e-3), "env_steps_per_s": world * ng * Kg / (float(tt[0]) * 1e-3),
                             "config": f"GroupedActionsObservations + FeatureVectorObservation, {WIDTH}x{HEIGHT}, gravity off, {ng} envs/GPU, "
                                       f"{A} placements x {A and gbase.layout.n_features} features per env-step, random legal placements, {Kg} steps (one event pair per step)"},
                 "rollout": {"placements_per_s": world * ng * A * Kr / (float(tt[1]) * 1e-3), "env_steps_per_s": world * ng * Kr / (float(tt[1]) * 1e-3),
                             "config": f"fused heuristic rollout, K = {Kr} steps per launch, {ng} envs/GPU, weights (-51, 76, -36, -18)"}}
        gbase.close()

    if rank == 0:
        peak, peak_src = peaks()
        bps, obs_bytes = bytes_per_step(layout, QUEUE, commit_frac)
        kernel_ms = ms / K     # rank-0 kernel: one k_step launch per step, back to back on the timed stream
        achieved = bps * n / (kernel_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"batched per-call step, {WIDTH}x{HEIGHT} board, padding 4, queue_size {QUEUE}, uniformly random actions (8 ids), "
                                   f"obs dict (board+mask+holder+queue) written every step, NEXT_STEP autoreset, Philox 7-bag",
                       "envs_per_gpu": n, "obs_bytes_per_env": obs_bytes, "l2_policy": "working set per step (%.2f GB) exceeds the 126 MB L2" % (bps * n / 1e9),
                       "parallelism": f"envs sharded over {world} GPU(s), no collective on the step path"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "tg::k_step_ws<10,20,uint32_t,0> (2 logic warps + 4 image warps per CTA)", "bytes_per_env_step": bps, "commit_frac": commit_frac,
                         "kernel_ms": kernel_ms, "peak_source": peak_src},
            "clocks": sampler.result(),
            "gpu_launches": K,
        }
        if e2e is not None:
            out["e2e"] = e2e
        if extra is not None:
            out["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline()
        tr = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(tr):
            try:
                tj = json.load(open(tr))
                out["roofline"]["traffic"] = tj.get("k_step_bytes_per_launch") if tj.get("envs") == n else None
                out["roofline"]["traffic_note"] = tj.get("source")
            except Exception:
                pass
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

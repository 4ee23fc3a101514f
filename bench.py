#!/usr/bin/env python
"""bench.py -- headline benchmark: batched per-call Tetris env step on B200 (BASELINE.json configs[1]), with the other
BASELINE configs as `extra` legs of the same JSON line.

    python bench.py --gpus N --steps K --warmup W            # our CUDA arm (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the unmodified reference on the host cores

A "step" = one `env.step(actions)` over ENVS_PER_GPU envs of the default 10x20 board, uniformly random actions over all 8
ids, observation dict (board + active mask + holder + queue) written every step, NEXT_STEP autoreset, device-native Philox
7-bag.  The population is first advanced into steady state (untimed), then W warm-up and K timed steps.  Prints ONE JSON
line (rank 0):
  value      device-timed whole-job env-steps/s (CUDA events, max over ranks), `roofline` = algorithmic HBM bytes / peak
  e2e        the same step through host buffers (`Tetris.step_host`, the reference's numpy-in / numpy-out convention):
             actions H2D from pinned memory, observation dict + 5-tuple in host arrays every step; `e2e.roofline` compares
             the host-memory traffic with the measured streaming-store ceiling of this host, `e2e.link` with the PCIe link
  extra      grouped placements/s (config 3), fused rollout at 2 M envs x K = 256 (config 4), wide board + RGB / CNN image
             (config 5), each with its own roofline block
  cpu_baseline  (N = 1) the C port of the reference on the host cores, DRAM-resident batch
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
STEADY_STEPS = 640                        # untimed steps before the timed window (several episode lengths under random actions)
WEIGHTS = (-51, 76, -36, -18)             # integer linear placement policy of the fused rollout


def workload_config(envs_per_gpu, world):
    """`config` of the JSON line -- identical for the CUDA arm and the reference arm (the CPU arms run a bounded sample of it,
    described in their `cpu_baseline.sample`)."""
    obs = 2 * (HEIGHT + 4) * (WIDTH + 8) + 16 + 16 * QUEUE
    return {"workload": f"batched per-call step, {WIDTH}x{HEIGHT} board, padding 4, queue_size {QUEUE}, uniformly random actions (8 ids), "
                        f"obs dict (board+mask+holder+queue) written every step, NEXT_STEP autoreset, 7-bag",
            "envs_per_gpu": envs_per_gpu, "obs_bytes_per_env": obs,
            "l2_policy": "working set per step (%.2f GB) exceeds the 126 MB L2" % ((obs + 262) * envs_per_gpu / 1e9),
            "parallelism": f"envs sharded over {world} GPU(s), no collective on the step path"}


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


def bytes_per_step(layout, queue, write_frac, reset_frac):
    """Algorithmic (compulsory) HBM bytes per env-step of OUR layout.  write_frac = share of env-steps that write
    their board record back (piece locked or env reset); the rng record is written when a bag is reshuffled (one lock in seven) or
    the env is reset."""
    ob = layout.obs_board_bytes
    obs = 2 * ob + 16 + 16 * queue
    read = layout.hot_stride + layout.board_stride + layout.rng_stride + 4
    write = layout.hot_stride + obs + 10 + write_frac * layout.board_stride + ((write_frac - reset_frac) / 7.0 + reset_frac) * layout.rng_stride
    return read + write, obs


# =================================================================================================================================
#  CPU arms
# =================================================================================================================================
def cpu_port(seconds=None, steps=20, warmup=3, n_envs=1 << 20):
    """The oracle port (oracle/tetris_oracle.c, OpenMP over the host cores) on a DRAM-resident 1 M-env sample of the workload."""
    from oracle.reference_arm import host_cores, port_throughput

    r = port_throughput(WIDTH, HEIGHT, QUEUE, n_envs, steps, warmup, cores=host_cores(), seconds=seconds)
    r.update({"unit": UNIT, "kind": "port",
              "sample": f"{n_envs} envs (DRAM-resident, 1/4 of the GPU arm's batch) x {r['steps']} vector steps ({r['seconds']:.1f} s), oracle/tetris_oracle.c "
                        f"via OpenMP x{r['cores']}, NEXT_STEP autoreset, numpy-exact 7-bag, obs dict written every step"})
    return r


def run_reference(args):
    """--impl reference: the UNMODIFIED reference (baseline/_ref or /root/reference) on all host cores, one worker process per
    core (B3); B1 / B2 and the C port as labelled context.  Falls back to the port when the reference cannot be imported."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import _refload
    from oracle.reference_arm import ReferenceWorkers, host_cores, reference_single_env, reference_sync_vector

    cores = host_cores()
    K, W = args.steps, max(args.warmup, 1)
    cfg = workload_config(args.envs, args.gpus)
    port = cpu_port(steps=max(4, min(K, 20)), warmup=2)
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg}
    if _refload.available():
        per_worker = 192          # ~0.1 s of reference work per vector step and core
        pool = ReferenceWorkers(WIDTH, HEIGHT, QUEUE, per_worker, cores)
        for _ in range(W):
            pool.step()
        t0 = time.perf_counter()
        for _ in range(K):
            pool.step()
        dt = time.perf_counter() - t0
        pool.close()
        v = pool.envs * K / dt
        b1 = reference_single_env(WIDTH, HEIGHT, QUEUE, seconds=2.0)
        b2 = reference_sync_vector(WIDTH, HEIGHT, QUEUE, m=64, seconds=2.0)
        sample = (f"{pool.envs} envs per vector step = {cores} worker processes x {per_worker} envs (bounded sample of the {args.envs}-env workload), "
                  f"unmodified tetris_gymnasium.envs.Tetris from {_refload.where()}, gymnasium stand-in (oracle/gymnasium_shim), NEXT_STEP autoreset, "
                  f"observation dicts stacked per worker")
        line.update({"value": v, "ms_per_step": 1e3 * dt / K,
                     "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
                     "reference_paths": {"B1_single_env_1_core": b1, "B2_sync_vector_64_envs_1_core": b2, "B3_one_worker_per_core": v, "unit": UNIT,
                                         "note": "B1 = examples/play_random.py loop without rendering; B2/B3 = SyncVectorEnv / AsyncVectorEnv stand-ins "
                                                 "(gymnasium is not installed)"},
                     "port": port})
    else:
        line.update({"value": port["value"], "ms_per_step": port["ms_per_step"],
                     "cpu_baseline": {"value": port["value"], "unit": UNIT, "cores": cores, "kind": "port", "sample": port["sample"]},
                     "note": "the reference package is not importable here (neither /root/reference nor baseline/_ref): the C port is timed instead"})
    line["e2e"] = {"value": line["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line))


# =================================================================================================================================
#  CUDA arm
# =================================================================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--steady", type=int, default=STEADY_STEPS, help="untimed steps that bring the population into steady state")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the grouped / rollout / wide-board legs")
    ap.add_argument("--no-probe", action="store_true", help="skip every torch kernel around the timed region (e.g. under ncu)")
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

    def allmax(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def allsum(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n, K, Wm = args.envs, args.steps, args.warmup
    start, stop = shard_range(world * n, rank, world)      # weak scaling: n envs per GPU, global ids keyed by shard
    env = Tetris(width=WIDTH, height=HEIGHT, gravity=True, queue_size=QUEUE, num_envs=stop - start, device=dev,
                 autoreset_mode="next_step", randomizer_mode="philox", env_id_offset=start)
    env.reset(seed=42)
    g = torch.Generator(device=dev)
    g.manual_seed(42 + rank)
    NA = 32                                                # distinct action vectors, cycled
    acts = torch.randint(0, 8, (NA, n), dtype=torch.int32, device=dev, generator=g)

    # ---- steady state: advance the population (untimed) until terminations are stationary -------------------------------------
    steady = {"untimed_steps": args.steady}
    term_sum, hist = torch.zeros((), dtype=torch.float64, device=dev), []
    for t in range(args.steady):
        _, _, term, _, _ = env.step(acts[t % NA])
        if not args.no_probe:
            term_sum += term.sum()
            if (t + 1) % 64 == 0:
                hist.append(term_sum.clone())
                term_sum.zero_()
    if hist:
        steady["terminated_per_env_step_by_64_step_block"] = [round(float(h) / (64 * n), 6) for h in hist]
    for t in range(Wm):
        env.step(acts[(args.steady + t) % NA])
    env.episode_stats(reset=True)
    snap = None if args.no_probe else (env._hot.clone(), env._brd.clone(), env._rng.clone())

    # ---- the timed window --------------------------------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    for t in range(K):
        env.step(acts[(args.steady + Wm + t) % NA])
    allreduce_episode_stats(env._stats)   # the only collective: episode statistics (4 doubles) over NCCL
    ev1.record()
    barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    ms_max = allmax(ms)[0]
    value = world * n * K / (ms_max * 1e-3)

    # ---- what the timed window did: replay it from the snapshot (the env is deterministic) and count, untimed -------------------
    write_frac, reset_frac, term_frac = 0.21, 0.004, 0.004
    if snap is not None:
        env._hot.copy_(snap[0]); env._brd.copy_(snap[1]); env._rng.copy_(snap[2])
        del snap
        # envs pending a NEXT_STEP reset when the window starts: bit 26 of hot word 0
        prev = ((env._hot.view(torch.int32)[::8] >> 26) & 1).to(torch.bool)
        w = torch.zeros(3, dtype=torch.float64, device=dev)
        for t in range(K):
            _, rew, term, _, _ = env.step(acts[(args.steady + Wm + t) % NA])
            w[0] += ((rew != 0) | term | prev).sum()      # board record written: piece locked (reward / game over) or env reset
            w[1] += prev.sum()
            w[2] += term.sum()
            prev = term.clone()
        write_frac, reset_frac, term_frac = [float(v) / (K * n) for v in w]
    steady.update({"board_write_frac_timed_window": write_frac, "reset_frac_timed_window": reset_frac, "terminated_frac_timed_window": term_frac})
    torch.cuda.synchronize()

    # ---- end to end through host buffers ---------------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = leg_e2e(env, acts, n, K, world, rank, dev, barrier, allmax, allsum)

    layout = env.layout
    extra = None
    if not args.no_extra:
        del env, acts
        torch.cuda.empty_cache()
        extra = leg_extra(n, world, rank, dev, barrier, allmax)

    if rank == 0:
        peak, peak_src = peaks()
        bps, obs_bytes = bytes_per_step(layout, QUEUE, write_frac, reset_frac)
        kernel_ms = ms / K     # rank-0 kernel: one k_step_ws launch per step, back to back on the timed stream
        achieved = bps * n / (kernel_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": workload_config(n, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "tg::k_step_ws<10,20,uint32_t,0,false> (2 logic warps + 4 image warps per CTA)", "bytes_per_env_step": bps,
                         "commit_frac": write_frac, "kernel_ms": kernel_ms, "peak_source": peak_src},
            "steady_state": steady,
            "clocks": sampler.result(),
            "gpu_launches": K,
        }
        if e2e is not None:
            out["e2e"] = e2e
        if extra is not None:
            out["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_port(seconds=10.0, warmup=2)
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        tr = os.path.join(ROOT, "profiles", "traffic_r02.json")
        if not os.path.exists(tr):
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


def leg_e2e(env, acts, n, K, world, rank, dev, barrier, allmax, allsum):
    """Tetris.step_host: actions from pinned host memory, observation dict + 5-tuple into pinned host arrays, every step."""
    import ctypes as C

    import torch
    from tetris_gymnasium_b200 import _lib

    lay = env.layout
    bufs = env.alloc_host_buffers(pinned=True)
    dict_bytes = sum(int(np.prod(bufs[k].shape)) for k in ("board", "active_tetromino_mask", "holder", "queue"))
    scal_bytes = sum(int(np.prod(bufs[k].shape)) * bufs[k].dtype.itemsize for k in ("reward", "terminated", "truncated", "lines_cleared"))
    Ke = min(K, 20)
    h_acts = torch.empty((Ke + 2, n), dtype=torch.int32, pin_memory=True)
    for t in range(Ke + 2):
        h_acts[t].copy_(acts[t % acts.shape[0]])
    h_np = h_acts.numpy()
    L = _lib.load()
    threads = int(os.environ.get("TG_HOST_THREADS", "0")) or max(1, len(os.sched_getaffinity(0)) // int(os.environ.get("LOCAL_WORLD_SIZE", "1")))

    # ceilings of this host, measured with every rank active at once: pinned D2H copy (no kernel) and streaming stores of the
    # expansion's thread count into the pinned observation buffer
    probe_bytes = min(1 << 30, bufs["board"].nbytes)
    d_probe = torch.empty(probe_bytes, dtype=torch.uint8, device=dev)
    h_probe = torch.from_numpy(bufs["board"].reshape(-1)[:probe_bytes])
    h_probe.copy_(d_probe)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        h_probe.copy_(d_probe, non_blocking=True)
    torch.cuda.synchronize()
    link_gbs = 3 * probe_bytes / (time.perf_counter() - t0) / 1e9
    barrier()
    bw = C.c_double()
    _lib.check(L.tg_host_membw(bufs["board"].ctypes.data, bufs["board"].nbytes, threads, 2, C.byref(bw)))
    store_gbs = bw.value
    barrier()
    del d_probe

    def run(mode, steps):
        for t in range(2):
            env.step_host(h_np[t], bufs, mode=mode)
        barrier()
        t0 = time.perf_counter()
        wait = exp = 0.0
        for t in range(steps):
            env.step_host(h_np[2 + t], bufs, mode=mode)
            st = env.host_stats()
            wait += st["wait_s"]; exp += st["expand_s"]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return allmax(dt)[0], wait / steps, exp / steps

    dt_c, wait_c, exp_c = run("compact", Ke)
    d2h_c = n * lay.host_record_bytes + (scal_bytes - n)      # packed link records + reward / terminated / lines (truncated is constant)
    Kd = min(Ke, 5)
    dt_d, _, _ = run("dma", Kd)
    d2h_d = dict_bytes + scal_bytes
    # the same loop when the policy lives on the GPU -- actions H2D, step, only the 5-tuple scalars D2H (the dict stays in HBM)
    h_rew = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h_term = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_act = torch.empty(n, dtype=torch.int32, device=dev)
    barrier()
    t1 = time.perf_counter()
    for t in range(Ke):
        d_act.copy_(h_acts[2 + t], non_blocking=True)
        _, rew, term, _, _ = env.step(d_act)
        h_rew.copy_(rew, non_blocking=True)
        h_term.copy_(term.view(torch.uint8), non_blocking=True)
        torch.cuda.synchronize()
    dt_s = allmax(time.perf_counter() - t1)[0]
    link_sum, store_sum = allsum(link_gbs, store_gbs)
    v = world * n * Ke / dt_c
    host_written = (dict_bytes + d2h_c + n) * Ke / dt_c / 1e9        # per GPU: dict (streaming stores) + packed records and scalars (DMA)
    return {"value": v, "unit": UNIT, "steps": Ke, "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": d2h_c,
            "mode": "compact", "host_threads_per_gpu": threads, "ms_per_step": 1e3 * dt_c / Ke,
            "note": "Tetris.step_host(mode='compact') = tg_step_host(TG_HOST_COMPACT): actions from pinned host memory; the step runs without the dict, packed "
                    "link records (hot words 0, 2, 3 + nibble id plane) and reward / terminated / lines cross PCIe chunk by chunk, and the library's host threads rebuild the full "
                    "observation dict in the caller's pinned arrays with streaming stores while later chunks are still in flight; every step delivers the same "
                    "bytes as the device-written dict (tests/test_gpu_host_step.py)",
            "breakdown_ms_per_step_rank0": {"waiting_for_device": 1e3 * wait_c, "expanding": 1e3 * exp_c},
            "roofline": {"bound": "host-memory-write", "achieved": host_written, "peak": store_gbs, "unit": "GB/s", "frac": host_written / store_gbs,
                         "bytes_written_to_host_memory_per_env_step": (dict_bytes + d2h_c + n) / n,
                         "peak_source": f"streaming stores of {threads} host threads into the pinned buffer, every rank at once (tg_host_membw); "
                                        f"sum over ranks {store_sum:.1f} GB/s"},
            "link": {"d2h_pinned_gb_s_per_gpu": link_gbs, "d2h_pinned_gb_s_sum_over_ranks": link_sum,
                     "compact_link_gb_s": (4 * n + d2h_c) * Ke / dt_c / 1e9, "frac_of_link": (4 * n + d2h_c) * Ke / dt_c / 1e9 / link_gbs,
                     "note": "pinned D2H copy of 1 GiB x 3, no kernel, all ranks concurrently"},
            "dma_mode": {"value": world * n * Kd / dt_d, "unit": UNIT, "steps": Kd, "d2h_bytes_per_step": d2h_d,
                         "link_gb_s": (4 * n + d2h_d) * Kd / dt_d / 1e9, "frac_of_link": (4 * n + d2h_d) * Kd / dt_d / 1e9 / link_gbs,
                         "note": "tg_step_host(TG_HOST_DMA): the dict is written on the device and DMA-copied (round-1 path)"},
            "obs_on_device": {"value": world * n * Ke / dt_s, "unit": UNIT, "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 5 * n,
                              "note": "same loop with the observation dict left in HBM (GPU-resident policy): actions H2D, step, reward + terminated D2H, synchronised every step"}}


def issue_block(rate_env_steps, profiles_and_envs, sm_mhz=1965.0, sms=148):
    """Issue-slot roofline of an integer-issue-bound leg: warp instructions per env-step (smsp__inst_executed.sum of the ncu
    captures under profiles/ divided by the env-steps of the captured launch) x the env-step rate measured HERE, against the
    issue peak of the chip (SMs x 4 schedulers x 1 warp instruction per clock at the maximum SM clock)."""
    per_step, src = 0.0, []
    for name, env_steps in profiles_and_envs:
        try:
            j = json.load(open(os.path.join(ROOT, "profiles", name)))
            per_step += float(j["smsp__inst_executed.sum"]["value"]) / env_steps
            src.append(f"profiles/{name}")
        except Exception:
            return None
    peak_issue = sms * 4 * sm_mhz * 1e6
    return {"bound": "issue", "warp_inst_per_env_step": per_step, "achieved": per_step * rate_env_steps / 1e9, "peak": peak_issue / 1e9,
            "unit": "G warp-inst/s", "frac": per_step * rate_env_steps / peak_issue, "source": " + ".join(src) + " (smsp__inst_executed.sum / env-steps of the captured launch)"}


def leg_extra(n, world, rank, dev, barrier, allmax):
    """The other BASELINE configs, short runs, each with a roofline block (HBM bytes are algorithmic, per our layout)."""
    import torch
    import torch.distributed as dist
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from tetris_gymnasium_b200.sharding import allreduce_episode_stats
    from tetris_gymnasium_b200.wrappers import CnnObservation, FeatureVectorObservation, GroupedActionsObservations, RgbObservation

    peak, _ = peaks()
    extra = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- config 3: GroupedActionsObservations + FeatureVectorObservation ------------------------------------------------------
    ng = min(n, 1 << 20)
    gbase = Tetris(width=WIDTH, height=HEIGHT, gravity=False, queue_size=4, num_envs=ng, device=dev, env_id_offset=rank * ng)
    genv = GroupedActionsObservations(gbase, observation_wrappers=[FeatureVectorObservation(gbase)])
    genv.reset(seed=42)
    Kg, tg_ms = 24, 0.0
    for t in range(Kg + 4):
        a = torch.multinomial(genv.legal_actions_mask.float() + 1e-9, 1).squeeze(1).to(torch.int32)   # random legal placement (untimed)
        ev0.record()
        genv.step(a)
        ev1.record()
        torch.cuda.synchronize()
        if t >= 4:
            tg_ms += ev0.elapsed_time(ev1)
    tg_ms = allmax(tg_ms)[0]
    lay = gbase.layout
    A, F = lay.n_placements, lay.n_features
    # placement step: state R + W (every env commits), action, legal mask R, 5-tuple; enumeration: state R, features + mask + info W
    g_bytes = 2 * (lay.hot_stride + lay.board_stride) + lay.rng_stride + 4 + A + 10 + (lay.hot_stride + lay.board_stride) + A * F + A + F
    g_rate = ng * Kg / (tg_ms * 1e-3)
    extra["grouped"] = {"placements_per_s": world * g_rate * A, "env_steps_per_s": world * g_rate,
                        "config": f"GroupedActionsObservations + FeatureVectorObservation, {WIDTH}x{HEIGHT}, gravity off, {ng} envs/GPU, "
                                  f"{A} placements x {F} features per env-step, random legal placements, {Kg} steps (one event pair per step)",
                        "roofline": {"bound": "hbm", "achieved": g_bytes * g_rate / 1e9, "peak": peak, "unit": "GB/s", "frac": g_bytes * g_rate / 1e9 / peak,
                                     "bytes_per_env_step": g_bytes, "kernels": "tg::k_step_ws<10,20,u32,2,false> + tg::k_grouped_feats_x<10,u32>",
                                     "note": "integer-issue bound, not HBM bound: see the `issue` block"}}
    extra["grouped"]["issue"] = issue_block(g_rate, [("r02_gfeats_x_steady.json", 1 << 20), ("r02_step_grouped_steady.json", 1 << 20)])
    gbase.close()
    del genv, gbase
    torch.cuda.empty_cache()

    # ---- config 4: fused K-step heuristic rollout, 2 M envs per GPU, K = 256, stats all-reduce inside the timed region ----------
    nr, Kr = min(2 * n, 1 << 21) if n >= (1 << 20) else n, 256
    rbase = Tetris(width=WIDTH, height=HEIGHT, gravity=False, queue_size=QUEUE, num_envs=nr, device=dev, env_id_offset=rank * nr)
    rbase.reset(seed=42)
    rbase.rollout(WEIGHTS, 16)
    rbase.episode_stats(reset=True)
    barrier()
    ev0.record()
    rbase.rollout(WEIGHTS, Kr)
    allreduce_episode_stats(rbase._stats)
    ev1.record()
    barrier()
    tr_ms = allmax(ev0.elapsed_time(ev1))[0]
    st = rbase._stats.cpu().tolist()
    lay = rbase.layout
    r_bytes = 2 * (lay.hot_stride + lay.board_stride + lay.rng_stride) / Kr
    r_rate = nr * Kr / (tr_ms * 1e-3)
    extra["rollout"] = {"placements_per_s": world * r_rate * lay.n_placements, "env_steps_per_s": world * r_rate,
                        "config": f"fused heuristic rollout, {WIDTH}x{HEIGHT}, holder + 7-bag, queue {QUEUE}, K = {Kr} steps per launch, {nr} envs/GPU "
                                  f"({world * nr} total), weights {WEIGHTS}, episode statistics all-reduced (NCCL) inside the timed region",
                        "episode_stats_all_ranks": {"episodes": st[0], "mean_return": st[1] / max(st[0], 1), "mean_length": st[2] / max(st[0], 1), "mean_lines": st[3] / max(st[0], 1)},
                        "roofline": {"bound": "hbm", "achieved": r_bytes * r_rate / 1e9, "peak": peak, "unit": "GB/s", "frac": r_bytes * r_rate / 1e9 / peak,
                                     "bytes_per_env_step": r_bytes, "kernels": "tg::k_rollout_x<10,u32>",
                                     "note": "state touches HBM once per K steps: integer-issue bound by design, see the `issue` block"}}
    extra["rollout"]["issue"] = issue_block(r_rate, [("r02_rollout.json", (1 << 20) * 64)])
    rbase.close()
    del rbase
    torch.cuda.empty_cache()

    # ---- config 5: wide board 20x40, queue 5, image observation (RGB for a CNN; and the fused 84x84 grey frame stack) ------------
    nw = min(n, 1 << 18)
    wbase = Tetris(width=20, height=40, queue_size=5, num_envs=nw, device=dev, env_id_offset=rank * nw)
    wenv = RgbObservation(wbase)
    wenv.reset(seed=42)
    gq = torch.Generator(device=dev)
    gq.manual_seed(7 + rank)
    wa = torch.randint(0, 8, (8, nw), dtype=torch.int32, device=dev, generator=gq)
    Kw = 24
    for t in range(64):
        wenv.step(wa[t % 8])
    barrier()
    ev0.record()
    for t in range(Kw):
        wenv.step(wa[t % 8])
    ev1.record()
    barrier()
    tw_ms = allmax(ev0.elapsed_time(ev1))[0]
    lay = wbase.layout
    img = lay.height_padded * lay.rgb_width * 3
    # dict-less step (state R, hot W, ~0.15 board W) + image kernel (state R, image W)
    w_bytes = 2 * (lay.hot_stride + lay.board_stride) + lay.rng_stride + lay.hot_stride + 0.15 * lay.board_stride + 14 + img
    w_rate = nw * Kw / (tw_ms * 1e-3)
    extra["wide_rgb"] = {"env_steps_per_s": world * w_rate, "image_bytes_per_env": img,
                         "config": f"20x40 board, queue 5, RgbObservation u8[{lay.height_padded},{lay.rgb_width},3], {nw} envs/GPU, random actions, {Kw} steps",
                         "roofline": {"bound": "hbm", "achieved": w_bytes * w_rate / 1e9, "peak": peak, "unit": "GB/s", "frac": w_bytes * w_rate / 1e9 / peak,
                                      "bytes_per_env_step": w_bytes, "kernels": "tg::k_step_ws<20,40,u64,0> (no dict) + tg::k_rgb"}}
    wbase.close()
    del wenv, wbase
    torch.cuda.empty_cache()
    nc = min(n, 1 << 16)
    cbase = Tetris(width=20, height=40, queue_size=5, num_envs=nc, device=dev, env_id_offset=rank * nc)
    cenv = CnnObservation(cbase, shape=(84, 84), stack_size=4, window=28, clip_reward=True)
    cenv.reset(seed=42)
    ca = torch.randint(0, 8, (8, nc), dtype=torch.int32, device=dev, generator=gq)
    for t in range(8):
        cenv.step(ca[t % 8])
    barrier()
    ev0.record()
    for t in range(Kw):
        cenv.step(ca[t % 8])
    ev1.record()
    barrier()
    tc_ms = allmax(ev0.elapsed_time(ev1))[0]
    c_bytes = 2 * (lay.hot_stride + lay.board_stride) + lay.rng_stride + lay.hot_stride + 0.15 * lay.board_stride + 14 + 84 * 84 * (1 + 2 * 3 / 28)
    c_rate = nc * Kw / (tc_ms * 1e-3)
    extra["wide_cnn"] = {"env_steps_per_s": world * c_rate,
                         "config": f"20x40 board, queue 5, fused CNN adapter (RGB -> 84x84 INTER_AREA -> grey -> 4-frame stack, clip reward), {nc} envs/GPU, {Kw} steps",
                         "roofline": {"bound": "hbm", "achieved": c_bytes * c_rate / 1e9, "peak": peak, "unit": "GB/s", "frac": c_bytes * c_rate / 1e9 / peak,
                                      "bytes_per_env_step": c_bytes, "kernels": "tg::k_step_ws<20,40,u64,0> (no dict) + tg::k_cnn_obs3",
                                      "note": "integer-issue bound (fixed-point resize + grey per output pixel), see the `issue` block"}}
    extra["wide_cnn"]["issue"] = issue_block(c_rate, [("r02_cnn.json", 1 << 16)])
    cbase.close()
    del cenv, cbase
    torch.cuda.empty_cache()

    # ---- functional facade: batched_step with the State in and out every call -----------------------------
    from tetris_gymnasium_b200.envs import tetris_fn as FN
    from tetris_gymnasium_b200.functional.core import EnvConfig
    from tetris_gymnasium_b200.functional.tetrominoes import TETROMINOES
    nf = min(n, 1 << 20)
    fcfg = EnvConfig(width=10, height=20, padding=4, queue_size=7)
    with torch.cuda.device(dev):
        keys = torch.stack([torch.arange(nf, device=dev) + rank * nf, torch.full((nf,), 42, device=dev)], dim=1)
        keys, fstate, _ = FN.batched_reset(TETROMINOES, keys, config=fcfg)
        fa = torch.randint(0, 7, (8, nf), dtype=torch.int32, device=dev, generator=gq)
        for t in range(8):
            fstate, _, _, _, _ = FN.batched_step(TETROMINOES, fstate, fa[t % 8], config=fcfg)
        barrier()
        ev0.record()
        for t in range(Kw):
            fstate, _, _, _, _ = FN.batched_step(TETROMINOES, fstate, fa[t % 8], config=fcfg)
        ev1.record()
        barrier()
    tf_ms = allmax(ev0.elapsed_time(ev1))[0]
    f_bytes = 2 * 24 * 18 + 200 + 2 * 4 * 16 + 4 + 9      # board i8 in + out, observation i8[20,10], scalars in + out, action, 5-tuple
    f_rate = nf * Kw / (tf_ms * 1e-3)
    extra["functional"] = {"env_steps_per_s": world * f_rate,
                           "config": f"functional facade batched_step (envs/tetris_fn.py), 10x20, queue 7, {nf} envs/GPU, State (int8 board + scalars) in and out every call, {Kw} steps",
                           "roofline": {"bound": "hbm", "achieved": f_bytes * f_rate / 1e9, "peak": peak, "unit": "GB/s", "frac": f_bytes * f_rate / 1e9 / peak,
                                        "bytes_per_env_step": f_bytes, "kernels": "tg::k_fn_step_tile<16>"}}
    return extra


if __name__ == "__main__":
    main()

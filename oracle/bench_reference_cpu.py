"""Time the UNMODIFIED reference (Max-We/Tetris-Gymnasium, imported from /root/reference through oracle/_refload.py) on the
host cores of the BUILD container -- the CPU baselines B1-B5 of BASELINE.md section 4.  Test / measurement infrastructure:
the reference is pure Python and cannot travel to the GPU box, so these context numbers are recorded here, next to the
core count, in profiles/r01_reference_cpu.json; `bench.py` times the C oracle port on the GPU box instead.

    python -m oracle.bench_reference_cpu [--seconds 4] [--out profiles/r01_reference_cpu.json]

gymnasium itself is not installed: SyncVectorEnv / AsyncVectorEnv are labelled stand-ins (an in-process loop over M envs;
one worker process per core, each stepping its share of the envs and returning the stacked observation through a pipe --
the work gymnasium.vector.AsyncVectorEnv does without shared memory).  JAX is absent: B6 is "not measurable".
"""
import argparse
import json
import multiprocessing as mp
import os
import time

import numpy as np

from . import _refload


def _loop(env, seconds, n_actions, rng, grouped=False):
    env.reset(seed=42)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        if grouped:
            mask = env.legal_actions_mask
            legal = np.flatnonzero(mask)
            a = int(rng.choice(legal)) if len(legal) else 0
        else:
            a = int(rng.integers(0, n_actions))
        _, _, term, _, _ = env.step(a)
        steps += 1
        if term:
            env.reset()
    return steps / (time.perf_counter() - t0)


def _worker(conn, n_envs, seconds, seed):
    R = _refload.load()
    envs = [R["make"]() for _ in range(n_envs)]
    for i, e in enumerate(envs):
        e.reset(seed=seed + i)
    rng = np.random.default_rng(seed)
    conn.send("ready")
    conn.recv()
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        obs = []
        for e in envs:
            o, _, term, _, _ = e.step(int(rng.integers(0, 8)))
            if term:
                o, _ = e.reset()
            obs.append(o["board"])
        conn.send(np.stack(obs))          # the observation crosses the process boundary every step, like AsyncVectorEnv
        conn.recv()
        steps += n_envs
    conn.send(("done", steps, time.perf_counter() - t0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    R = _refload.load()
    rng = np.random.default_rng(42)
    cores = len(os.sched_getaffinity(0))
    rows = []

    def out(name, value, unit, note, used=1):
        rows.append({"baseline": name, "value": value, "unit": unit, "cores_used": used, "host_cores": cores, "note": note})
        print(f"{name:58s} {value:12.1f} {unit}  ({note})", flush=True)

    env = R["make"](gravity=True)
    out("B1 reference Tetris, single env, 10x20, gravity on", _loop(env, args.seconds, 8, rng), "env-steps/s", "random actions, reset on game over")
    for M in (8, 64):
        envs = [R["make"]() for _ in range(M)]
        for i, e in enumerate(envs):
            e.reset(seed=42 + i)
        steps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < args.seconds:
            obs = []
            for e in envs:
                o, _, term, _, _ = e.step(int(rng.integers(0, 8)))
                if term:
                    o, _ = e.reset()
                obs.append(o)
            _ = {k: np.stack([o[k] for o in obs]) for k in obs[0]}
            steps += M
        out(f"B2 SyncVectorEnv stand-in (in-process loop), M = {M}", steps / (time.perf_counter() - t0), "env-steps/s", "observation dicts stacked every step")
    per = 8
    pipes, procs = [], []
    for w in range(cores):
        a, b = mp.Pipe()
        pr = mp.Process(target=_worker, args=(b, per, args.seconds, 1000 * (w + 1)), daemon=True)
        pr.start()
        pipes.append(a); procs.append(pr)
    for a in pipes:
        a.recv()
    for a in pipes:
        a.send("go")
    total, tmax, alive = 0, 0.0, set(range(cores))
    while alive:
        for w in list(alive):
            m = pipes[w].recv()
            if isinstance(m, tuple):
                total += m[1]; tmax = max(tmax, m[2]); alive.discard(w)
            else:
                pipes[w].send("next")
    for pr in procs:
        pr.join(timeout=5)
    out(f"B3 AsyncVectorEnv stand-in ({cores} worker processes x {per} envs)", total / tmax, "env-steps/s", "board observation piped to the parent every step", used=cores)
    base = R["make"](gravity=False)
    g = R["GroupedActionsObservations"](base, observation_wrappers=[R["FeatureVectorObservation"](base)])
    v = _loop(g, args.seconds, 40, rng, grouped=True)
    out("B4 reference GroupedActionsObservations + FeatureVectorObservation", v, "env-steps/s", f"= {40 * v:.0f} placements/s, random legal placements")
    wide = R["RgbObservation"](R["make"](width=20, height=40, queue_size=5))
    out("B5 reference wide board 20x40 + RgbObservation", _loop(wide, args.seconds, 8, rng), "env-steps/s", "random actions")
    rows.append({"baseline": "B6 jit(vmap(step)) functional env on the JAX CPU backend", "value": None, "unit": "env-steps/s",
                 "note": "not measurable: jax / chex are not installed in this image"})
    if args.out:
        json.dump({"where": "build container (no GPU)", "python": os.sys.version.split()[0], "numpy": np.__version__, "rows": rows},
                  open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

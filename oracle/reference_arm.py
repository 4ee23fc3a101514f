"""CPU arms of bench.py (MEASUREMENT INFRASTRUCTURE, not product code).

(1) The UNMODIFIED reference (`tetris_gymnasium.envs.Tetris`, imported through oracle/_refload.py from /root/reference or
    from the pip --target install in baseline/_ref): B1 single env like examples/play_random.py:7-13, B2 a
    SyncVectorEnv-style in-process loop, B3 one worker process per host core (what gymnasium.vector.AsyncVectorEnv does;
    gymnasium itself is not installed in this image, so the vector envs are labelled stand-ins that do the same work:
    step every env, NEXT_STEP autoreset, stack the observation dicts into per-worker arrays).
(2) The C port of the reference (oracle/tetris_oracle.c, OpenMP over all host cores) on a DRAM-resident batch.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from . import _refload


def host_cores():
    """Every core this process may run on (torchrun exports OMP_NUM_THREADS=1 to its workers, so the OpenMP default is not
    trusted; thread / worker counts are passed explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ---- the C port ---------------------------------------------------------------------------------------------------------------
def port_throughput(width, height, queue, n_envs, steps, warmup, cores=None, seconds=None):
    """env-steps/s of the oracle port: `n_envs` envs (numpy-exact 7-bag, NEXT_STEP autoreset, obs dict written every step),
    `steps` timed vector steps after `warmup` (or as many as fit `seconds`)."""
    from .tetris_oracle import OracleVec

    cores = cores or host_cores()
    vec = OracleVec(n_envs, bulk=True, width=width, height=height, gravity=True, queue_size=queue)
    vec.seed_all(1 + np.arange(n_envs, dtype=np.uint64))
    vec.reset_all(cores)
    rng = np.random.default_rng(42)
    acts = rng.integers(0, 8, size=(8, n_envs)).astype(np.int32)
    for t in range(warmup):
        vec.step(acts[t % 8], nthreads=cores)
    t0 = time.perf_counter()
    done = 0
    while True:
        vec.step(acts[done % 8], nthreads=cores)
        done += 1
        if (seconds is None and done >= steps) or (seconds is not None and time.perf_counter() - t0 > seconds):
            break
    dt = time.perf_counter() - t0
    return {"value": n_envs * done / dt, "cores": cores, "steps": done, "seconds": dt, "envs": n_envs,
            "ms_per_step": 1e3 * dt / done}


# ---- the unmodified reference ---------------------------------------------------------------------------------------------------
def _make(R, width, height, queue):
    return R["make"](width=width, height=height, gravity=True, queue_size=queue)


def _vec_step(envs, pending, actions, out):
    """One SyncVectorEnv-style step (NEXT_STEP autoreset) over `envs`, observation dicts stacked into `out`."""
    for i, e in enumerate(envs):
        if pending[i]:
            o, _ = e.reset()
            pending[i] = False
        else:
            o, _, term, _, _ = e.step(int(actions[i]))
            pending[i] = term
        for k in out:
            out[k][i] = o[k]


def _worker(conn, cfg, n_envs, seed):
    R = _refload.load()
    envs = [_make(R, *cfg) for _ in range(n_envs)]
    o = None
    for i, e in enumerate(envs):
        o, _ = e.reset(seed=seed + i)
    out = {k: np.empty((n_envs,) + v.shape, v.dtype) for k, v in o.items()}
    pending = [False] * n_envs
    rng = np.random.default_rng(seed)
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg == "stop":
            break
        _vec_step(envs, pending, rng.integers(0, 8, size=n_envs), out)
        conn.send("done")


class ReferenceWorkers:
    """One worker process per core, each stepping `per_worker` unmodified reference envs per vector step."""

    def __init__(self, width, height, queue, per_worker, cores=None):
        self.cores = cores or host_cores()
        self.per_worker = per_worker
        ctx = mp.get_context("fork")
        self.pipes, self.procs = [], []
        for w in range(self.cores):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(b, (width, height, queue), per_worker, 1000 * (w + 1)), daemon=True)
            p.start()
            self.pipes.append(a)
            self.procs.append(p)
        for a in self.pipes:
            assert a.recv() == "ready"

    @property
    def envs(self):
        return self.cores * self.per_worker

    def step(self):
        for a in self.pipes:
            a.send("step")
        for a in self.pipes:
            a.recv()

    def close(self):
        for a in self.pipes:
            try:
                a.send("stop")
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)


def reference_single_env(width, height, queue, seconds=2.0):
    """B1: examples/play_random.py:7-13 without the rendering -- reset(seed=42), random actions until game over, repeat."""
    R = _refload.load()
    env = _make(R, width, height, queue)
    env.reset(seed=42)
    rng = np.random.default_rng(42)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        _, _, term, _, _ = env.step(int(rng.integers(0, 8)))
        steps += 1
        if term:
            env.reset()
    return steps / (time.perf_counter() - t0)


def reference_sync_vector(width, height, queue, m=64, seconds=2.0):
    """B2: SyncVectorEnv-style loop over m envs in one process."""
    R = _refload.load()
    envs = [_make(R, width, height, queue) for _ in range(m)]
    o = None
    for i, e in enumerate(envs):
        o, _ = e.reset(seed=42 + i)
    out = {k: np.empty((m,) + v.shape, v.dtype) for k, v in o.items()}
    pending = [False] * m
    rng = np.random.default_rng(42)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        _vec_step(envs, pending, rng.integers(0, 8, size=m), out)
        steps += m
    return steps / (time.perf_counter() - t0)

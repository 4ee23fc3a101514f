"""GPU: the host-buffer step (tg_step_host) -- the compact path (packed records over PCIe + host-side dict expansion) against
the full-dict DMA path and against the device-resident step, bit for bit; ordering after asynchronous work on the caller's stream."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KEYS = ("board", "active_tetromino_mask", "holder", "queue", "reward", "terminated", "truncated", "lines_cleared")


def _same(a, b, ctx):
    for k in KEYS:
        if not np.array_equal(a[k], b[k]):
            bad = np.flatnonzero((a[k] != b[k]).reshape(len(a[k]), -1).any(1))
            raise AssertionError(f"{ctx}: {k} differs for envs {bad[:8]} (of {len(bad)})")


@pytest.mark.parametrize("cfg,n,pinned,threads", [
    (dict(width=10, height=20, queue_size=7), 1000, True, 0),
    (dict(width=10, height=20, queue_size=7), 40000, True, 3),
    (dict(width=10, height=20, queue_size=4, gravity=False), 4097, False, 1),
    (dict(width=20, height=40, queue_size=5), 3000, True, 2),
    (dict(width=7, height=9, queue_size=3), 2500, True, 0),
    (dict(width=13, height=21, queue_size=1), 777, False, 2),
    (dict(width=24, height=12, queue_size=16), 1500, True, 0),
])
def test_compact_host_step_equals_dma_and_device(cfg, n, pinned, threads):
    from tetris_gymnasium_b200.envs.tetris import Tetris
    from gpu_util import np_

    rng = np.random.default_rng(n)
    mk = lambda: Tetris(num_envs=n, autoreset_mode="next_step", randomizer_mode="philox", **cfg)  # noqa: E731
    dev, dma, cpt = mk(), mk(), mk()
    for e in (dev, dma, cpt):
        e.reset(seed=11)
    cpt.set_host_threads(threads)
    b_dma, b_cpt = dma.alloc_host_buffers(pinned=pinned), cpt.alloc_host_buffers(pinned=pinned)
    for t in range(60):
        a = rng.choice([0, 1, 2, 3, 4, 5, 5, 6, 7], size=n).astype(np.int32)
        obs, r, term, trunc, info = dev.step(torch.from_numpy(a))
        o1 = dma.step_host(a, b_dma, mode="dma")
        o2 = cpt.step_host(a, b_cpt, mode="compact")
        _same(o1, o2, f"dma vs compact, t={t}")
        for k in ("board", "active_tetromino_mask", "holder", "queue"):
            assert np.array_equal(np_(obs[k]), o2[k]), (t, k)
        assert np.array_equal(np_(r), o2["reward"]) and np.array_equal(np_(term), o2["terminated"].astype(bool))
        assert np.array_equal(np_(info["lines_cleared"]), o2["lines_cleared"]) and not o2["truncated"].any()
    st = cpt.host_stats()
    assert st["chunks"] >= 1 and st["total_s"] > 0
    for e in (dev, dma, cpt):
        e.close()


def test_compact_host_step_many_chunks(monkeypatch):
    """Small chunks and a 2-slot ring: the chunk pipeline (slot reuse, ragged last chunk) must not change a byte."""
    from tetris_gymnasium_b200.envs.tetris import Tetris

    monkeypatch.setenv("TG_HOST_CHUNK", "1024")
    monkeypatch.setenv("TG_HOST_RING", "2")
    n = 20000 + 37
    rng = np.random.default_rng(2)
    a_env = Tetris(num_envs=n, queue_size=7)
    b_env = Tetris(num_envs=n, queue_size=7)
    a_env.reset(seed=3); b_env.reset(seed=3)
    ba, bb = a_env.alloc_host_buffers(), b_env.alloc_host_buffers()
    for t in range(30):
        a = rng.integers(0, 8, size=n).astype(np.int32)
        _same(a_env.step_host(a, ba, mode="dma"), b_env.step_host(a, bb, mode="compact"), f"t={t}")
    assert b_env.host_stats()["chunks"] == (n + 1023) // 1024


def test_host_step_is_ordered_after_the_callers_stream():
    """tg_step_host runs on internal streams; it must wait for a reset / set_state the caller enqueued on its own stream
    (a large reset is still running when the host call starts)."""
    from tetris_gymnasium_b200.envs.tetris import Tetris

    n = 1 << 20
    e1, e2 = Tetris(num_envs=n, queue_size=7), Tetris(num_envs=n, queue_size=7)
    bufs1, bufs2 = e1.alloc_host_buffers(), e2.alloc_host_buffers()
    a = np.full(n, 5, np.int32)
    for rep in range(3):
        e1.reset(seed=100 + rep)
        torch.cuda.synchronize()
        o1 = e1.step_host(a, bufs1, mode="compact")
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            e2.reset(seed=100 + rep)                       # asynchronous, not synchronised
            o2 = e2.step_host(a, bufs2, mode="compact")
        _same(o1, o2, f"rep {rep}")

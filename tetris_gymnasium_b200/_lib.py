"""ctypes binding of libtetris_b200.so (the C ABI in include/tetris_b200.h).

This is the only place the Python host touches native code.  The library is plain CUDA behind
`extern "C"`; torch is used by the callers for device memory and streams only.  There is no CPU
fallback: if the shared library is missing and cannot be built, importing raises.
"""
import ctypes as C
import os

from . import _build

TG_OK = 0
AUTORESET = {"disabled": 0, "next_step": 1, "same_step": 2}
RNG = {"philox": 0, "sequence": 1, "numpy": 2}
RANDOMIZER = {"bag": 0, "true": 1}
TG_SCALARS = 8
TG_OPT_TERMINATE_ON_ILLEGAL, TG_OPT_HOST_THREADS = 1, 2
HOST_MODE = {"dma": 0, "compact": 1}
TG_VERSION = 2


class TgConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("queue_size", C.c_int32), ("gravity", C.c_int32),
        ("autoreset", C.c_int32), ("rng_mode", C.c_int32),
        ("action_map", C.c_int32 * 8),
        ("terminate_on_illegal", C.c_int32), ("randomizer", C.c_int32),
        ("reward_alife", C.c_double), ("reward_clear_line", C.c_double),
        ("reward_game_over", C.c_double), ("reward_invalid_action", C.c_double),
        ("seq_len", C.c_int64), ("env_id_offset", C.c_uint64),
        ("holder_size", C.c_int32), ("n_pieces", C.c_int32),
        ("piece_n", C.c_uint8 * 8), ("piece_matrix", (C.c_uint8 * 16) * 7), ("piece_color", (C.c_uint8 * 3) * 7), ("reserved1", C.c_uint8 * 3),
    ]


class TgLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "width_padded", "height_padded", "hot_stride", "board_stride", "rng_stride", "obs_board_bytes",
        "obs_holder_bytes", "obs_queue_bytes", "n_placements", "n_features", "rgb_width", "host_record_bytes")]


class TgState(C.Structure):
    _fields_ = [("hot", C.c_void_p), ("board", C.c_void_p), ("rng", C.c_void_p), ("piece_seq", C.c_void_p)]


class TgObs(C.Structure):
    _fields_ = [("board", C.c_void_p), ("mask", C.c_void_p), ("holder", C.c_void_p), ("queue", C.c_void_p)]


class TgStepOut(C.Structure):
    _fields_ = [("reward", C.c_void_p), ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("lines", C.c_void_p)]


EXPORTS = (
    "tg_create tg_destroy tg_get_layout tg_last_error tg_version tg_reset tg_seed_numpy tg_step tg_step_host "
    "tg_features tg_render_rgb tg_grouped_observe tg_grouped_step tg_rollout tg_get_state tg_set_state tg_debug_set_rollout_trace tg_fn_step tg_cnn_observe "
    "tg_set_host_threads tg_host_stats tg_host_expand tg_seed_numpy_seeds tg_host_membw tg_step_n tg_set_option"
).split()

_LIB = None


def lib_path():
    return _build.LIB


def load():
    """Load (building in-tree with nvcc if needed) the native library.  Raises when impossible."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if _build.needs_build():
        try:
            _build.build()
        except Exception as exc:  # a stale-but-present .so is still usable (e.g. on the GPU box without nvcc)
            if not os.path.exists(path):
                raise RuntimeError(f"libtetris_b200.so is missing and could not be built: {exc}") from exc
    L = C.CDLL(path)
    vp, i64 = C.c_void_p, C.c_int64
    L.tg_version.restype = C.c_int
    L.tg_last_error.restype = C.c_char_p
    L.tg_last_error.argtypes = [vp]
    L.tg_create.argtypes = [C.POINTER(TgConfig), C.c_int, C.POINTER(vp)]
    L.tg_destroy.argtypes = [vp]
    L.tg_set_option.argtypes = [vp, C.c_int32, i64]
    L.tg_get_layout.argtypes = [vp, C.POINTER(TgLayout)]
    L.tg_reset.argtypes = [vp, TgState, i64, vp, vp, TgObs, vp]
    L.tg_seed_numpy.argtypes = [vp, TgState, i64, vp, vp, vp]
    L.tg_seed_numpy_seeds.argtypes = [vp, TgState, i64, vp, vp, vp]
    L.tg_step.argtypes = [vp, TgState, i64, vp, TgObs, TgStepOut, vp, vp]
    L.tg_step_n.argtypes = [vp, TgState, i64, C.c_int32, vp, TgObs, i64, TgStepOut, i64, vp, vp]
    L.tg_step_host.argtypes = [vp, TgState, i64, vp, TgObs, TgStepOut, C.c_int32, vp]
    L.tg_set_host_threads.argtypes = [vp, C.c_int32]
    L.tg_host_stats.argtypes = [vp, C.POINTER(C.c_double)]
    L.tg_host_membw.argtypes = [vp, i64, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    L.tg_host_expand.argtypes = [C.POINTER(TgConfig), i64, vp, vp, TgObs, C.c_int32]
    L.tg_features.argtypes = [vp, TgState, i64, vp, vp]
    L.tg_render_rgb.argtypes = [vp, TgState, i64, vp, vp]
    L.tg_cnn_observe.argtypes = [vp, TgState, i64, C.c_int32, C.c_int32, vp, i64, vp, C.c_int32, vp]
    L.tg_grouped_observe.argtypes = [vp, TgState, i64, vp, vp, vp, vp]
    L.tg_grouped_step.argtypes = [vp, TgState, i64, vp, vp, vp, vp, vp, TgObs, TgStepOut, vp, vp]
    L.tg_rollout.argtypes = [vp, TgState, i64, C.POINTER(C.c_int32), C.c_int32, vp, vp]
    L.tg_debug_set_rollout_trace.argtypes = [vp, vp]
    L.tg_fn_step.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp]
    L.tg_get_state.argtypes = [vp, TgState, i64, vp, vp, vp]
    L.tg_set_state.argtypes = [vp, TgState, i64, vp, vp, vp, vp]
    for name in EXPORTS:
        if name not in ("tg_last_error",):
            getattr(L, name).restype = C.c_int
    _LIB = L
    return L


class TgError(RuntimeError):
    pass


def check(rc, handle=None):
    if rc != TG_OK:
        msg = load().tg_last_error(handle)
        raise TgError(f"libtetris_b200 error {rc}: {msg.decode() if msg else '?'}")

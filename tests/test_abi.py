"""CPU: the C-ABI library builds, loads and exports every symbol include/tetris_b200.h declares.
No compute calls (no GPU here); without a device tg_create must fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "tetris_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tg_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from tetris_gymnasium_b200 import _lib

    L = _lib.load()
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/tetris_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert L.tg_version() == 2


def test_struct_layouts_match_header():
    from tetris_gymnasium_b200 import _lib

    # sizes implied by the C declarations (natural alignment)
    assert C.sizeof(_lib.TgConfig) == 6 * 4 + 8 * 4 + 2 * 4 + 4 * 8 + 8 + 8 + 2 * 4 + 8 + 7 * 16 + 7 * 3 + 3
    assert C.sizeof(_lib.TgLayout) == 12 * 4
    assert C.sizeof(_lib.TgState) == 4 * 8 and C.sizeof(_lib.TgObs) == 4 * 8 and C.sizeof(_lib.TgStepOut) == 4 * 8


def test_no_cpu_fallback():
    import torch

    from tetris_gymnasium_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.load()
    cfg = _lib.TgConfig()
    cfg.width, cfg.height, cfg.queue_size = 10, 20, 4
    for i in range(8):
        cfg.action_map[i] = i
    h = C.c_void_p()
    assert L.tg_create(C.byref(cfg), 0, C.byref(h)) == 3  # TG_ERR_CUDA
    assert b"no CPU fallback" in L.tg_last_error(None)
    from tetris_gymnasium_b200.envs.tetris import Tetris

    with pytest.raises(RuntimeError):
        Tetris(num_envs=4)


def test_config_validation_precedes_device_probe():
    from tetris_gymnasium_b200 import _lib

    L = _lib.load()
    cfg = _lib.TgConfig()
    cfg.width, cfg.height, cfg.queue_size = 30, 20, 4  # padded row would not fit 32 bits
    h = C.c_void_p()
    assert L.tg_create(C.byref(cfg), 0, C.byref(h)) == 1  # TG_ERR_CONFIG
    assert b"width" in L.tg_last_error(None)


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (checked textually)."""
    pkg = os.path.join(ROOT, "tetris_gymnasium_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f

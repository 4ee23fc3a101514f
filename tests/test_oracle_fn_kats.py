"""CPU: the reference's functional-env tests (tests/test_functional/test_core/*.py, test_env/*.py, test_queue.py), restated in
tests/fn_kats.py, against the numpy oracle of the functional env.  tests/test_gpu_fn_kats.py runs the same cases on tg_fn_step."""
import pytest

import fn_kats


@pytest.mark.parametrize("case", fn_kats.ALL, ids=lambda f: f.__name__)
def test_fn_kat_on_oracle(case):
    trace = case(lambda first, gravity: fn_kats.OracleDriver(first, gravity))
    assert len(trace) > 0

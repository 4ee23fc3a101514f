"""CPU oracle for the Tetris hot path -- TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package.  The product package `tetris_gymnasium_b200` never does.
"""

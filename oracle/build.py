"""Compile oracle/tetris_oracle.c -> oracle/liboracle.so (gcc, OpenMP)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tetris_oracle.c")
LIB = os.path.join(HERE, "liboracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-Wall", "-Wextra", "-std=gnu11", "-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

"""Build libtetris_b200.so in-tree with nvcc for sm_100a (no torch involved: the library is plain CUDA + C ABI)."""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtetris_b200.so")
BUILD = os.path.join(HERE, "build")


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = glob.glob(os.path.join(CSRC, "*")) + [os.path.join(os.path.dirname(HERE), "include", "tetris_b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


def _host_objects(verbose=False):
    """The host-side dict expansion of the compact host step (csrc/tg_host_expand*.cpp/.inc): plain g++, the loop compiled once
    per instruction-set variant (picked at run time from the CPU the library is loaded on)."""
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    objs = []
    common = [cxx, "-O3", "-std=c++17", "-fPIC", "-pthread", "-c", "-I", CSRC]
    jobs = [("tg_host_expand.o", ["tg_host_expand.cpp"], [])]
    for suffix, flags in (("base", []), ("avx2", ["-mavx2", "-mbmi2"]),
                          ("avx512", ["-mavx512f", "-mavx512bw", "-mavx512vl", "-mavx512vbmi", "-mavx2", "-mbmi2"])):
        jobs.append((f"tg_host_expand_{suffix}.o", ["-x", "c++", "tg_host_expand_impl.inc"], [f"-DTGH_SUFFIX={suffix}"] + flags))
    for out, src, flags in jobs:
        o = os.path.join(BUILD, out)
        cmd = common + flags + [a if a.startswith("-") or a == "c++" else os.path.join(CSRC, a) for a in src] + ["-o", o]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        objs.append(o)
    return objs


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libtetris_b200.so (there is no CPU fallback)")
    os.makedirs(BUILD, exist_ok=True)
    # one rank per GPU may import the package at once: build under a lock, into a temporary file, then rename atomically
    import fcntl
    with open(os.path.join(BUILD, ".lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        if not force and not needs_build():
            return LIB
        objs = _host_objects(verbose)
        tmp = LIB + f".tmp{os.getpid()}"
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-shared", "--use_fast_math", "-Xptxas", "-v" if verbose else "-O3",
               "-o", tmp, os.path.join(CSRC, "tg_api.cu")] + objs + ["-lcudart", "-lpthread"] + os.environ.get("TG_NVCC_FLAGS", "").split()
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, LIB)
        if verbose:
            print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))

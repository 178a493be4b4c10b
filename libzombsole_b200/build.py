"""Compiles csrc/zs_b200.cu into csrc/libzs_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libzs_b200.so")
SOURCES = ["zs_b200.cu"]
HEADERS = ["zs_device.cuh", "zs_world.cuh", "zs_obs.cuh", os.path.join("..", "..", "include", "zs_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--fmad=false", "-cudart", "static",
              "-Xcompiler", "-fopenmp", "-lgomp"]  # (OpenMP: the host-side expansion of compact observation records)


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_native(force=False, verbose=False, extra=()):
    """Build the shared library in-tree; returns its path."""
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + ["-o", LIB] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    env = dict(os.environ)
    # a stray CC/CXX in the environment may point at a compiler driver without libstdc++ headers
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(cmd, cwd=CSRC, env=env)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_native(force=True, verbose="-v" in sys.argv))

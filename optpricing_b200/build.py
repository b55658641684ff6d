"""
Build liboptmc.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m optpricing_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "optmc.cu")
OUT = os.path.join(HERE, "liboptmc.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _sources():
    d = os.path.join(HERE, "csrc")
    inc = os.path.join(os.path.dirname(HERE), "include")
    return [os.path.join(d, f) for f in os.listdir(d)] + [os.path.join(inc, "optmc.h")]


def up_to_date() -> bool:
    return os.path.exists(OUT) and all(
        os.path.getmtime(OUT) >= os.path.getmtime(s) for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc, *NVCC_FLAGS, "-o", OUT, SRC]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

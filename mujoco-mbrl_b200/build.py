"""Builds the C-ABI shared library in-tree with nvcc for sm_100a.

    python mujoco-mbrl_b200/build.py [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmbrl_b200.so")
SOURCES = [os.path.join(CSRC, "mbrl_b200.cu")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr", "--expt-extended-lambda",
    "-ldl",
]


def _deps():
    out = list(SOURCES)
    for name in os.listdir(CSRC):
        if name.endswith((".cuh", ".h")):
            out.append(os.path.join(CSRC, name))
    out.append(os.path.join(os.path.dirname(HERE), "include", "mbrl_b200.h"))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    """Compile libmbrl_b200.so if missing or stale.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmbrl_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + SOURCES
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libmbrl_b200.so (see output above)")
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

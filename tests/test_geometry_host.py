"""CPU test: the kernels' shared-memory / TMEM geometry functions (host code in the .cuh files) hold
their invariants over a sweep of model shapes -- alignment of every operand/tile/barrier offset, TMEM
column budget, shared-memory budget, padding large enough for the zero-padded straight-line loops."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_geometry_sweep(tmp_path):
    exe = str(tmp_path / "geometry_check")
    src = os.path.join(ROOT, "tests", "host", "geometry_check.cu")
    build = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--expt-relaxed-constexpr",
                            "--expt-extended-lambda", "-I", os.path.join(ROOT, "include"), "-o", exe, src],
                           capture_output=True, text=True, timeout=900)
    assert build.returncode == 0, build.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout[-3000:]
    assert "failures 0" in run.stdout

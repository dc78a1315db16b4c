"""CPU test of bench.py's reference arm (`--impl reference`): the JSON line carries every key the
driver's contract names.  (The B200 arm needs a GPU; its line is checked by the same key list in
profiles/bench_r2_1gpu.json, which a GPU run of bench.py produced.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _check_common(line):
    assert BASE_KEYS <= set(line), sorted(BASE_KEYS - set(line))
    assert line["unit"] == "candidate-steps/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    assert line["cpu_baseline"]["kind"] in ("port", "reference")


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cartpole",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    _check_common(line)
    assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] == line["cpu_baseline"]["value"]


@pytest.mark.parametrize("name", ["bench_r1_1gpu.json", "bench_r2_1gpu.json"])
def test_committed_b200_line_has_the_contract_keys(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    _check_common(line)
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-9
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert line["gpu_launches"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["e2e"]["value"] != line["value"]

"""The bench.py contract pieces that run without a GPU: the reference arm (`--impl reference`, the reference's own solver on the
host cores) prints one JSON line with the keys the driver reads, and the sparse workloads answer `unavailable` for that arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsmallk_ref.so")


def _last_json(out):
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert lines, out[-2000:]
    return json.loads(lines[-1])


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built on this machine")
def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--size", "1500"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = _last_json(r.stdout)
    assert line["impl"] == "reference" and line["metric"] == "nmf_outer_iterations_per_second" and line["unit"] == "iter/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1000.0) < 1e-6 * 1000.0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


@pytest.mark.parametrize("workload", ["c3", "c4"])
def test_sparse_workloads_have_no_reference_arm(workload):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload], capture_output=True, text=True,
                       timeout=120, cwd=ROOT)
    assert r.returncode == 0
    line = _last_json(r.stdout)
    assert line["impl"] == "reference" and "unavailable" in line

"""bench.py prints ONE JSON line with the contract's keys (driver-facing); checked on a small run."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_bench_line_has_the_contract_keys():
    d = _run("--steps", "40", "--warmup", "3", "--envs-per-gpu", "8192", "--sets", "2", "--e2e-steps", "3")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "gpu_launches", "clocks", "roofline", "e2e", "cpu_baseline"):
        assert key in d, key
    assert d["steps"] == 40 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["gpu_launches"] == 40 + 1 and d["value"] > 0      # 40 step kernels + the fold of the one logging step (i = 0)
    assert d["stats_collective"]["in_timed_region"] is True
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]            # PCIe inside the timed region: e2e cannot repeat the device number
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--envs-per-gpu", "4096", "--sets", "2")
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]
    ours = _run("--steps", "8", "--warmup", "3", "--envs-per-gpu", "4096", "--sets", "2", "--no-extras")
    assert ours["config"] == d["config"] and ours["metric"] == d["metric"] and ours["unit"] == d["unit"]   # same_config for the driver

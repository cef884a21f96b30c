"""bench.py's JSON contract on the CPU: the reference arm (`--impl reference`: the oracle port timed on the host cores -- the one
bench leg that may execute oracle/) prints ONE line with the driver's keys, on the same `config`, `metric` and `unit` as the GPU
arm, and reports the steps it actually ran.  (The GPU arm's line is exercised by the driver on a B200.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["config"] == bench.workload_config(1)                 # same workload description as the GPU arm
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - d["sample_envs_per_step"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "512" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


import pytest


@pytest.mark.gpu
def test_gpu_arm_line():
    """The GPU arm at N = 1 with a short region: every key the driver reads, internally consistent."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "6", "--warmup", "3", "--no-extra"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["n_gpus"] == 1 and d["steps"] == 6 and d["warmup"] == 3
    assert d["config"] == bench.workload_config(1) and d["dtype"] == "bf16" and d["data"] == "synthetic" and d["scaling"] == "weak"
    E = d["config"]["envs_per_gpu"]
    assert abs(d["value"] - E / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    e = d["e2e"]
    assert 0 < e["value"] <= 1.05 * d["value"] and e["h2d_bytes_per_step"] == E * 13 * 32 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("tensor", "hbm") and r["unit"] in ("TFLOP/s", "GB/s") and 0 < r["frac"] < 1.0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and (r["traffic"] is None or r["traffic"] > 0)
    assert max(k["share_of_step"] for k in r["kernels"].values()) == r["share_of_step"]      # the kernel with the largest share
    assert d["gpu_launches"] == 6 * 27                       # 1 env step + patch-embed GEMM + 12 x 2 fused blocks + final LN/pool
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] > 0 and cb["cores"] >= 1
    c = d["clocks"]
    assert c["sm_max_mhz"] >= c["sm_mhz"] > 0 and isinstance(c["reasons"], list)

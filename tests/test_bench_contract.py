"""The bench line contract (driver-facing JSON keys) checked on the committed lines under profiles/ and on the
argument handling of bench.py — no GPU, no timing."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"}


def _line(path):
    lines = [l for l in path.read_text().splitlines() if l.startswith("{")]
    assert len(lines) == 1, f"{path.name}: expected exactly one JSON line, found {len(lines)}"
    return json.loads(lines[0])


@pytest.mark.parametrize("name", ["r02_bench_n1.json", "r02_bench_n2.json", "r02_bench_n4.json", "r02_bench_n8.json",
                                  "r02_bench_strong_n2.json", "r02_bench_strong_n4.json"])
def test_committed_bench_lines_follow_the_contract(name):
    d = _line(ROOT / "profiles" / name)
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert d["metric"].split(" ")[0].replace("->", "→") in base["metric"]
    assert d["unit"] == "samples/s" and d["higher_is_better"] is True and d["dtype"] == "bf16"
    assert d["data"] == "synthetic" and d["vs_baseline"] is None      # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and d["config"]["global_batch"] == (
        64 if d["scaling"] == "strong" else 64 * d["n_gpus"])
    assert d["warmup"] >= 3 and d["steps"] >= 1
    # value = whole-job samples / max-over-ranks time
    assert d["value"] == pytest.approx(d["config"]["global_batch"] / (d["ms_per_step"] / 1e3), rel=1e-6)
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0.5 * d["value"] < e["value"] <= 1.02 * d["value"]           # a measured number, not a copy of `value`
    assert e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-6) and 0.0 < r["frac"] < 1.0
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] <= c["sm_max_mhz"] and not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(
        c["reasons"])
    if d["n_gpus"] == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]


def test_bench_requires_a_gpu_for_the_b200_arm_and_knows_its_flags():
    """No CPU fallback: without a CUDA device the B200 arm exits with a message instead of measuring something else."""
    import torch

    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--help"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--scaling"):
        assert flag in out.stdout
    if torch.cuda.is_available():
        pytest.skip("the refusal is only observable without a GPU")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)

"""The bench.py JSON contract, checked on the committed evidence (profiles/r01_bench_*.json): every key the
driver reads is present and self-consistent.  Runs on CPU; the numbers themselves come from B200 runs."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_n*.json")))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_bench_line_has_the_contract_keys(path):
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "clips/s" and d["higher_is_better"] is True and d["dtype"] == "bf16" and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    # value is the whole-job aggregate: global batch / step time
    assert abs(d["value"] - d["config"]["global_batch"] / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == "clips/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] > 0
    if d["n_gpus"] == 1:
        r, c = d["roofline"], d["cpu_baseline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert r["traffic"] is None or r["traffic"] > 0
        assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_reference_arm.json")))
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0

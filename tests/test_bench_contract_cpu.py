"""The bench.py JSON contract, checked on the committed evidence (profiles/r0*_bench_*.json): every key the
driver reads is present and self-consistent.  Runs on CPU; the numbers themselves come from B200 runs."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r0*_bench_n*.json")) +
               glob.glob(os.path.join(ROOT, "profiles", "r02_bench_config*.json")))


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


@pytest.mark.parametrize("name", ["r01_bench_reference_arm.json", "r02_bench_reference_arm.json"])
def test_reference_arm_line(name):
    d = json.load(open(os.path.join(ROOT, "profiles", name)))
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_round2_line_carries_the_same_gpu_torch_baseline_and_traffic():
    """Round 2: the default line times the reference's own ops on the same B200 and reads the dominant kernel's DRAM
    traffic from the committed ncu launch list (profiles/roofline_traffic.json, tools/roofline_traffic.py)."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")))
    t = d["torch_b200"]
    assert t["eager"]["value"] > 0 and t["compiled"]["value"] > 0 and t["eager"]["micro_batch"] == 64
    assert d["value"] > 10 * t["compiled"]["value"]
    r = d["roofline"]
    traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    assert r["kernel"] in traffic and traffic[r["kernel"]]["dram_bytes_per_launch"] > 0
    assert 0.5 < r["traffic"] / r["alg_bytes_per_launch"] <= 1.05       # no re-reads beyond the algorithmic bytes
    assert r["frac"] <= r["frac_of_rw_bound"] <= 1.0

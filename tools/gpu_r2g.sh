# kernel + model tests after the (kT,3,3) tiled forward / stream state work; first bench lines of configs 2, 4, 5
mkdir -p gpurun_out; rm -f gpurun_out/*.jsonl
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -k "dwconv or stream" > gpurun_out/pytest_r2g_dw.txt 2>&1; tail -5 gpurun_out/pytest_r2g_dw.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rf -s --deselect tests/test_kernels_gpu.py::test_dwconv_fwd_dgrad_wgrad > gpurun_out/pytest_r2g.txt 2>&1
tail -6 gpurun_out/pytest_r2g.txt
for c in 2 5 4; do
  PB_BENCH_DETAIL=gpurun_out/detail_cfg$c.txt timeout 900 python bench.py --config $c --steps 3 --warmup 3 --torch-compile-budget 0 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err || tail -5 gpurun_out/bench_cfg$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_cfg$c.json")); print("config $c:", d["metric"], d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("torch_b200") or {}).get("eager"), d["cpu_baseline"]["value"])
    for n,v in list(d["kernels"].items())[:8]: print("   %-24s %7.2f ms %5d x %7.0f GB/s" % (n, v["ms_per_step"], v["launches_per_step"], v["GBps"]))
except Exception as e: print("config $c failed", e)
PY
done

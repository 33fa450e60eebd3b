mkdir -p gpurun_out; rm -f gpurun_out/*.jsonl
python tools/bw_probe.py > gpurun_out/bw_probe.txt 2>&1; cat gpurun_out/bw_probe.txt
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short -rf -s -k "stream" > gpurun_out/pytest_r2h.txt 2>&1; tail -4 gpurun_out/pytest_r2h.txt; grep "stream bf16" gpurun_out/pytest_r2h.txt
PB_BENCH_DETAIL=gpurun_out/detail_cfg4.txt timeout 900 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err || tail -5 gpurun_out/bench_cfg4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_cfg4.json")); print("config 4:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"])
for n,v in list(d["kernels"].items())[:6]: print("   %-24s %7.2f ms %5d x %7.0f GB/s" % (n, v["ms_per_step"], v["launches_per_step"], v["GBps"]))
PY

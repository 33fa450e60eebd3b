# GPU regression + bench in one gpurun call: kernel and model parity tests, then the default bench with the
# per-kernel table.  usage: bash tools/gpu_check.sh [pytest-args]
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short "$@" 2>&1 | tail -3 | tee gpurun_out/pytest_tail.txt
PB_BENCH_DETAIL=gpurun_out/detail.txt timeout 400 python bench.py --steps 4 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json; d=json.load(open("gpurun_out/bench.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for n,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:18]: print("%-24s %7.2f ms %5d x %7.0f GB/s" % (n, v["ms_per_step"], v["launches_per_step"], v["GBps"]))
PY

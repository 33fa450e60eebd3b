mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_gemm_tc_gpu.py tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -k "eval or folded or stem" > gpurun_out/pytest_r3a.txt 2>&1; tail -6 gpurun_out/pytest_r3a.txt | cut -c1-250
grep "eval logits vs" gpurun_out/pytest_r3a.txt | head
timeout 900 python bench.py --config 2 --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err || tail -5 gpurun_out/bench_cfg2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_cfg2.json")); print("cfg2:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
for n,v in list(d["kernels"].items())[:6]: print("   %-24s %7.2f ms %5d x %7.0f GB/s" % (n, v["ms_per_step"], v["launches_per_step"], v["GBps"]))
PY
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_r3a.json 2> gpurun_out/bench_r3a.err || tail -5 gpurun_out/bench_r3a.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r3a.json")); print("cfg3:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
PY

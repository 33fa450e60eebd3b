mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -x -k "bn or norm or batch" > gpurun_out/pytest_r2y.txt 2>&1; tail -5 gpurun_out/pytest_r2y.txt | cut -c1-250
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_blocks_bf16_gpu.py tests/test_graph_gpu.py -m gpu -q --tb=short -rf -x > gpurun_out/pytest_r2y2.txt 2>&1; tail -5 gpurun_out/pytest_r2y2.txt | cut -c1-250
PB_BENCH_DETAIL=gpurun_out/detail_r2y.txt timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_r2y.json 2> gpurun_out/bench_r2y.err || tail -5 gpurun_out/bench_r2y.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2y.json")); print("cfg3:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
for n,v in list(d["kernels"].items())[:10]: print("   %-24s %7.2f ms %5d x %7.0f GB/s" % (n, v["ms_per_step"], v["launches_per_step"], v["GBps"]))
PY

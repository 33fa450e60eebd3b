mkdir -p gpurun_out
timeout 300 python tools/bn_bench.py > gpurun_out/bn_bench.txt 2>&1; cat gpurun_out/bn_bench.txt | tail -15
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_graph_gpu.py -m gpu -q --tb=short -rf -x -k "bn or norm or graph" > gpurun_out/pytest_r2z.txt 2>&1; tail -3 gpurun_out/pytest_r2z.txt | cut -c1-250
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err || tail -5 gpurun_out/bench_r2z.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2z.json")); print("cfg3:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
PY

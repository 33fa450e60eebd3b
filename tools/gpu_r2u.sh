mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -x > gpurun_out/pytest_r2u.txt 2>&1; tail -3 gpurun_out/pytest_r2u.txt | cut -c1-250
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err || tail -5 gpurun_out/bench_r2u.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2u.json")); print("cfg3:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
PY
TMA_BENCH_NPROD=1 timeout 300 tools/_build/tma_bench > gpurun_out/tma_bench_nprod.txt 2>&1; cat gpurun_out/tma_bench_nprod.txt

mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -x > gpurun_out/pytest_r2t.txt 2>&1; tail -5 gpurun_out/pytest_r2t.txt | cut -c1-250
bash tools/gpu_r2s.sh 2>&1 | grep -v hsig
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err || tail -5 gpurun_out/bench_r2t.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2t.json")); print("cfg3:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
PY

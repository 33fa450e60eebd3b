mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -rf -x > gpurun_out/pytest_r2v.txt 2>&1; tail -3 gpurun_out/pytest_r2v.txt | cut -c1-250
S="1,702464,40,240 1,702464,240,40 1,225792,112,672 1,225792,672,112 1,6422528,16,64 1,1204224,72,24 1,175616,80,184 1,59584,160,960 1,59584,960,160"
echo "== nprod auto"; timeout 300 python tools/pw_bench.py gemm $S 2>&1 | tail -9
echo "== nprod 1"; PB_GEMM_NPROD=1 timeout 300 python tools/pw_bench.py gemm $S 2>&1 | tail -9
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_r2v.json 2> gpurun_out/bench_r2v.err || tail -5 gpurun_out/bench_r2v.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2v.json")); print("cfg3:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
PY

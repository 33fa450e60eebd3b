mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -rf -x > gpurun_out/pytest_r3d.txt 2>&1; tail -3 gpurun_out/pytest_r3d.txt | cut -c1-250
S="1,225792,112,672 1,200704,112,480 1,59584,160,960 1,47040,160,960 1,175616,80,480 1,702464,40,240"
echo "== n-resident"; timeout 300 python tools/pw_bench.py gemm $S 2>&1 | tail -6
echo "== off"; PB_GEMM_NO_NRES=1 timeout 300 python tools/pw_bench.py gemm $S 2>&1 | tail -6
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_r3d.json 2> gpurun_out/bench_r3d.err || tail -5 gpurun_out/bench_r3d.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r3d.json")); print("cfg3:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["kernels"]["pb_pw_gemm_tc"]["ms_per_step"])
PY

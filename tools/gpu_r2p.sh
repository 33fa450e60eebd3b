mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_graph_gpu.py tests/test_blocks_bf16_gpu.py -m gpu -q --tb=short -rf -x > gpurun_out/pytest_r2p.txt 2>&1; tail -8 gpurun_out/pytest_r2p.txt | cut -c1-250
PB_BENCH_DETAIL=gpurun_out/detail_r2p.txt timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err || tail -5 gpurun_out/bench_r2p.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2p.json")); print("cfg3:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
for n,v in list(d["kernels"].items())[:8]: print("   %-24s %7.2f ms %5d x %7.0f GB/s" % (n, v["ms_per_step"], v["launches_per_step"], v["GBps"]))
PY
grep "pb_pw_gemm_tc|64,0" gpurun_out/detail_r2p.txt | head; grep "pb_pw_gemm_tc" gpurun_out/detail_r2p.txt | sort -k1 -n -r | head -24

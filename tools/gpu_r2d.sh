# round-2 GPU call: mma depthwise v6 (8-channel items, 16-byte loads, conflict-free context mapping)
mkdir -p gpurun_out; rm -f gpurun_out/*.jsonl
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -x -k dwconv > gpurun_out/pytest_r2d_dw.txt 2>&1; tail -3 gpurun_out/pytest_r2d_dw.txt
timeout 300 python tools/dw_bench.py > gpurun_out/dwbench_mma.txt 2>&1; cat gpurun_out/dwbench_mma.txt
timeout 120 python tools/dw_bench.py 4.5 --only fwd --reps 2 > gpurun_out/plain_45.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dw_s1_mma -c 2 -o gpurun_out/mma6_45 python tools/dw_bench.py 4.5 --only fwd --reps 2 > gpurun_out/ncu_45.log 2>&1
timeout 120 python tools/dw_bench.py 3.2 --only fwd --reps 2 > gpurun_out/plain_32.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dw_s1_mma -c 2 -o gpurun_out/mma6_32 python tools/dw_bench.py 3.2 --only fwd --reps 2 > gpurun_out/ncu_32.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rf -s --deselect tests/test_kernels_gpu.py::test_dwconv_fwd_dgrad_wgrad > gpurun_out/pytest_r2d.txt 2>&1
tail -4 gpurun_out/pytest_r2d.txt
PB_BENCH_DETAIL=gpurun_out/detail_r2d.txt timeout 600 python bench.py --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2d.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"].get("depthwise_conv3d_frac"))
PY

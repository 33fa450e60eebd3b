# round-2 second GPU call: TMA micro-benchmark (fixed), depthwise micro-benchmark old vs mma kernels, kernel + block + model tests, bench
mkdir -p gpurun_out; rm -f gpurun_out/*.jsonl
tools/_build/tma_bench > gpurun_out/tma_bench.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -x -k dwconv > gpurun_out/pytest_r2b_dw.txt 2>&1; tail -3 gpurun_out/pytest_r2b_dw.txt
PB_DW_MMA=0 timeout 300 python tools/dw_bench.py > gpurun_out/dwbench_old.txt 2>&1
timeout 300 python tools/dw_bench.py > gpurun_out/dwbench_mma.txt 2>&1; tail -4 gpurun_out/dwbench_old.txt gpurun_out/dwbench_mma.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rf -s --deselect tests/test_kernels_gpu.py::test_dwconv_fwd_dgrad_wgrad > gpurun_out/pytest_r2b.txt 2>&1
tail -5 gpurun_out/pytest_r2b.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r2b.txt 2>&1; tail -3 gpurun_out/smoke_r2b.txt
PB_BENCH_DETAIL=gpurun_out/detail_r2b.txt timeout 600 python bench.py --steps 4 --warmup 3 --torch-compile-budget 0 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2b.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"].get("depthwise_conv3d_frac"))
PY

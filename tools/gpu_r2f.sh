mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -x -k dwconv > gpurun_out/pytest_r2f_dw.txt 2>&1; tail -2 gpurun_out/pytest_r2f_dw.txt
timeout 300 python tools/dw_bench.py --only fwd > gpurun_out/dwbench_mma.txt 2>&1; cat gpurun_out/dwbench_mma.txt

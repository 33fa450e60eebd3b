mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -x -k dwconv > gpurun_out/pytest_r2e_dw.txt 2>&1; tail -3 gpurun_out/pytest_r2e_dw.txt
timeout 300 python tools/dw_bench.py > gpurun_out/dwbench_mma.txt 2>&1; cat gpurun_out/dwbench_mma.txt
timeout 120 python tools/dw_bench.py 4.5 --only fwd --reps 2 > gpurun_out/plain_45.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dw_s1_mma -c 2 -o gpurun_out/mma7_45 python tools/dw_bench.py 4.5 --only fwd --reps 2 > gpurun_out/ncu_45.log 2>&1

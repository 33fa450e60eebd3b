mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -x -k "stem" > gpurun_out/pytest_r2k.txt 2>&1; tail -15 gpurun_out/pytest_r2k.txt
timeout 300 python tools/stem_bench.py > gpurun_out/stem_bench.txt 2>&1; cat gpurun_out/stem_bench.txt
PB_STEM_GATHER=1 timeout 300 python tools/stem_bench.py 2>&1 | sed 's/^/gather: /'

timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -k "stem" > gpurun_out/pytest_r2k.txt 2>&1; tail -5 gpurun_out/pytest_r2k.txt | cut -c1-250
timeout 300 python tools/stem_bench.py > gpurun_out/stem_bench.txt 2>&1; cat gpurun_out/stem_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_tma -c 2 -o gpurun_out/r02_stem_tma2 -f python tools/stem_bench.py --reps 1 > gpurun_out/ncu_stem.log 2>&1; tail -2 gpurun_out/ncu_stem.log

mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=line -x -k "stem_tma_path and 64x64 and mobilenet" > gpurun_out/sanitizer_stem.txt 2>&1; grep -v "^=========     at\|^=========         Host\|^=========     Host" gpurun_out/sanitizer_stem.txt | head -60

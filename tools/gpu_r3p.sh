timeout 300 python tools/pw_bench.py wgrad 2>&1 | tail -16
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -rf -x -k wgrad 2>&1 | tail -2

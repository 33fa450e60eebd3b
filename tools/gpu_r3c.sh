S="64,931,960,160 64,735,960,160 64,539,672,160 64,3528,672,112 64,3136,480,112 64,10976,120,40 1,225792,112,672 1,702464,40,240 1,6422528,16,64 1,175616,184,80"
echo "== nprod auto"; timeout 300 python tools/pw_bench.py wgrad $S 2>&1 | tail -10
echo "== nprod 1"; PB_WGRAD_NPROD=1 timeout 300 python tools/pw_bench.py wgrad $S 2>&1 | tail -10
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -rf -x -k wgrad 2>&1 | tail -2

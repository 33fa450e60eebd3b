S="1,125440,184,80 1,150528,184,80 1,175616,240,80 1,100352,240,80 1,702464,120,40 1,1204224,72,24 1,1204224,64,24 1,702464,240,40 1,301056,72,40"
echo "== wave fit"; timeout 300 python tools/pw_bench.py gemm $S 2>&1 | tail -9
echo "== off"; PB_GEMM_NO_WAVE_FIT=1 timeout 300 python tools/pw_bench.py gemm $S 2>&1 | tail -9
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -rf -x 2>&1 | tail -2

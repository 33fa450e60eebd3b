echo "== new"; timeout 300 python tools/gemm_stats_bench.py 2>&1 | tail -9
cp picklebot_b200/libpicklebot_b200.so /tmp/lib_new.so; cp tools/_build/lib_old.so picklebot_b200/libpicklebot_b200.so
echo "== old"; timeout 300 python tools/gemm_stats_bench.py 2>&1 | tail -9
cp /tmp/lib_new.so picklebot_b200/libpicklebot_b200.so
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -rf -x -k stat 2>&1 | tail -2

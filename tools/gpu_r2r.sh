mkdir -p gpurun_out
echo "== new"; timeout 300 python tools/small_bench.py 2>&1 | tail -12
cp picklebot_b200/libpicklebot_b200.so /tmp/lib_new.so; cp tools/_build/lib_old.so picklebot_b200/libpicklebot_b200.so
echo "== old"; timeout 300 python tools/small_bench.py 2>&1 | tail -12
cp /tmp/lib_new.so picklebot_b200/libpicklebot_b200.so

mkdir -p gpurun_out
timeout 300 python tools/stem_bench.py --reps 2 > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_tma -c 2 -o gpurun_out/r02_stem_tma -f python tools/stem_bench.py --reps 1 > gpurun_out/ncu_stem.log 2>&1; tail -3 gpurun_out/ncu_stem.log

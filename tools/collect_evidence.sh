# Round evidence from HEAD on one B200 (run under gpurun): tests, smoke, bench lines of every config, the ncu launch
# list of one eager micro-batch (shares + DRAM traffic per kernel family) and full ncu captures of the dominant kernels.
mkdir -p gpurun_out; rm -f gpurun_out/*.jsonl
timeout 2400 python -m pytest tests -m gpu -q --tb=short -rf -s > gpurun_out/r02_pytest.txt 2>&1; tail -4 gpurun_out/r02_pytest.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.txt 2>&1; tail -2 gpurun_out/r02_smoke.txt
python tools/bw_probe.py > gpurun_out/r02_bw_probe.txt 2>&1
timeout 300 python tools/dw_bench.py > gpurun_out/r02_dw_microbench.txt 2>&1
PB_DW_MMA=1 timeout 300 python tools/dw_bench.py > gpurun_out/r02_dw_microbench_mma_forced.txt 2>&1
timeout 300 python tools/bn_bench.py > gpurun_out/r02_bn_microbench.txt 2>&1
PB_BN_REGISTER_KERNELS=1 timeout 300 python tools/bn_bench.py > gpurun_out/r02_bn_microbench_register_kernels.txt 2>&1
timeout 300 python tools/stem_bench.py > gpurun_out/r02_stem_microbench.txt 2>&1
timeout 300 python tools/pw_bench.py wgrad > gpurun_out/r02_pw_wgrad_microbench.txt 2>&1
timeout 300 python tools/graph_timeline.py > gpurun_out/r02_graph_timeline.txt 2>/dev/null
PB_BENCH_DETAIL=gpurun_out/r02_kernel_detail_per_layer.txt timeout 1500 python bench.py --steps 4 --warmup 3 --torch-compile-budget ${COMPILE_BUDGET:-240} > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
for c in 2 4 5; do timeout 900 python bench.py --config $c --steps 3 --warmup 3 --torch-compile-budget 0 > gpurun_out/r02_bench_config$c.json 2> gpurun_out/r02_bench_config$c.err; done
timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-torch-b200 > gpurun_out/plain_range.log 2>&1 && \
PB_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/r02_ncu_launches_microbatch.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-torch-b200 > gpurun_out/ncu_range.log 2>&1
python tools/pw_bench.py gemm 1,702464,40,240 --reps 2 > gpurun_out/plain_pw.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -f -o gpurun_out/r02_gemm_tc_40_240 python tools/pw_bench.py gemm 1,702464,40,240 --reps 2 > gpurun_out/ncu_pw.log 2>&1
python tools/stem_bench.py --reps 2 > gpurun_out/plain_stem.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stem_tma -c 2 -f -o gpurun_out/r02_stem_tma_final python tools/stem_bench.py --reps 1 > gpurun_out/ncu_stem.log 2>&1
python - <<'PY'
import json
for f in ("r02_bench_n1", "r02_bench_config2", "r02_bench_config4", "r02_bench_config5"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("depthwise_conv3d_frac"))
    except Exception as e: print(f, "failed", e)
PY

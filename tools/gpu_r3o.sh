mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_blocks_bf16_gpu.py tests/test_graph_gpu.py -m gpu -q --tb=short -rf -x 2>&1 | tail -3 | cut -c1-200
PB_BENCH_DETAIL=gpurun_out/detail_r3o.txt timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('cfg3:', d['value'], d['ms_per_step'], d['roofline']['frac'], k['pb_pw_gemm_tc']['ms_per_step'], k.get('pb_fold_rows_bf16',{}).get('ms_per_step'))"
grep "fold_\|pb_pw_gemm_tc|1,1,59584,160,960\|pb_pw_gemm_tc|64,64" gpurun_out/detail_r3o.txt | head -24

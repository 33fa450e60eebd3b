timeout 300 python tools/se_dgrad_bench.py 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -rf -x 2>&1 | tail -3 | cut -c1-200
timeout 900 python -m pytest tests/test_blocks_bf16_gpu.py tests/test_models_gpu.py tests/test_graph_gpu.py -m gpu -q --tb=short -rf -x 2>&1 | tail -2 | cut -c1-200
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('cfg3:', d['value'], d['ms_per_step'], d['roofline']['frac'], k['pb_pw_gemm_tc']['ms_per_step'])"
timeout 900 python bench.py --config 2 --steps 4 --warmup 3 --no-torch-b200 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2:', d['value'], d['ms_per_step'])"

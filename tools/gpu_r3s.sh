timeout 300 python tools/bn_bench.py 2>&1 | tail -14
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q --tb=short -rf -x -k "bn or norm or train_step" 2>&1 | tail -2 | cut -c1-200
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('cfg3:', d['value'], d['ms_per_step'], k['pb_bn_act_fwd']['ms_per_step'], k['pb_bn_act_bwd_reduce']['ms_per_step'], k['pb_bn_act_bwd_apply']['ms_per_step'])"

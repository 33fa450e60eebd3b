timeout 300 python tools/bn_bench.py 2>&1 | tail -14
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -rf -x -k "bn or norm" 2>&1 | tail -2
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3:', d['value'], d['ms_per_step'])"

mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q --tb=short -rf -x > gpurun_out/pytest_r3g.txt 2>&1; tail -3 gpurun_out/pytest_r3g.txt | cut -c1-200
for c in 3 5; do timeout 900 python bench.py --config $c --steps 4 --warmup 3 --no-torch-b200 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('cfg$c:', d['value'], d['ms_per_step'], d['roofline']['frac'], 'colstats', k.get('pb_colstats',{}).get('ms_per_step'), k.get('pb_colstats',{}).get('GBps'))"; done

mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rf -x 2>&1 | tail -3 | cut -c1-200
for v in kernel memset kernel memset; do
if [ $v = memset ]; then export PB_ZERO_MEMSET=1; else unset PB_ZERO_MEMSET; fi
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('zero $v:', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done

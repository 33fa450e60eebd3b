timeout 1500 python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -3 | cut -c1-200
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('cfg3:', d['value'], d['ms_per_step'], d['roofline']['frac'], k['pb_pw_wgrad_tc']['ms_per_step'])"

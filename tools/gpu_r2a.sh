# round-2 first GPU call: TMA micro-benchmark, full GPU test-suite, smoke, bench with the torch-on-B200 legs
mkdir -p gpurun_out; rm -f gpurun_out/*.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
tools/_build/tma_bench > gpurun_out/tma_bench.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rf -s > gpurun_out/pytest_r2a.txt 2>&1
tail -5 gpurun_out/pytest_r2a.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r2a.txt 2>&1; tail -3 gpurun_out/smoke_r2a.txt
PB_BENCH_DETAIL=gpurun_out/detail_r2a.txt timeout 1300 python bench.py --steps 4 --warmup 3 --torch-compile-budget 700 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2a.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"]); print(json.dumps(d.get("torch_b200"))[:1500])
PY

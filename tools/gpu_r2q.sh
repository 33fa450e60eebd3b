mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_models_gpu.py -m gpu -q --tb=short -rf -x > gpurun_out/pytest_r2q.txt 2>&1; tail -5 gpurun_out/pytest_r2q.txt | cut -c1-250
timeout 600 python tools/graph_timeline.py > gpurun_out/graph_timeline.txt 2> gpurun_out/graph_timeline.err || tail -5 gpurun_out/graph_timeline.err; head -75 gpurun_out/graph_timeline.txt
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err || tail -5 gpurun_out/bench_r2q.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2q.json")); print("cfg3:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
PY

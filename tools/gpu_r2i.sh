mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -q --tb=short -rf -k "stream" > gpurun_out/pytest_r2i.txt 2>&1; tail -4 gpurun_out/pytest_r2i.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err || tail -20 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n2.json")); print("N=2:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("gradient_exchange"))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 tools/dp_check.py > gpurun_out/dp_check.txt 2>&1; tail -5 gpurun_out/dp_check.txt

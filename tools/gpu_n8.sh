mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err || tail -20 gpurun_out/bench_n8.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n8.json")); print("N=8:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("gradient_exchange"))
PY

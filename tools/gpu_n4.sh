mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err || tail -20 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json")); print("N=$N:", d["value"], d["ms_per_step"], d["e2e"]["value"])
PY

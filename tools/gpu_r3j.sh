mkdir -p gpurun_out
timeout 600 python bench.py --no-graphs --steps 2 --warmup 3 --no-torch-b200 --no-cpu 2>gpurun_out/nograph.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1 GPU eager:', d['value'], d['ms_per_step'])" || tail -5 gpurun_out/nograph.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dp_check.py 2>&1 | grep "dp_check" | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err || tail -10 gpurun_out/bench_n2.err
python -c "import json; d=json.load(open('gpurun_out/bench_n2.json')); print('N=2:', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 2 --warmup 3 --no-graphs 2> gpurun_out/bench_n2e.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=2 eager:', d['value'], d['ms_per_step'])" || tail -10 gpurun_out/bench_n2e.err

mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 > gpurun_out/bench_r3b$i.json 2> gpurun_out/bench_r3b.err || tail -5 gpurun_out/bench_r3b.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r3b$i.json")); print("cfg3 run $i:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernels"]["pb_pw_gemm_tc"]["ms_per_step"], d["kernels"]["pb_bn_act_bwd_apply"]["ms_per_step"])
PY
done
timeout 900 python bench.py --config 2 --steps 4 --warmup 3 --no-torch-b200 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2:', d['value'], d['ms_per_step'])"
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -rf -x 2>&1 | tail -2

mkdir -p gpurun_out
for v in on off on off; do
if [ $v = off ]; then export PB_GEMM_NO_NRES=1; else unset PB_GEMM_NO_NRES; fi
timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 --no-cpu > gpurun_out/bench_r3e_$v.json 2> gpurun_out/bench_r3e.err || tail -5 gpurun_out/bench_r3e.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r3e_$v.json")); print("nres $v:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["kernels"]["pb_pw_gemm_tc"]["ms_per_step"])
PY
done

mkdir -p gpurun_out
for cap in 0 300000 150000; do
PB_BN_ATOMIC_CAP=$cap PB_BENCH_DETAIL=gpurun_out/detail_cap$cap.txt timeout 900 python bench.py --steps 3 --warmup 3 --no-torch-b200 > gpurun_out/bench_cap$cap.json 2> gpurun_out/bench_cap.err || tail -5 gpurun_out/bench_cap.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_cap$cap.json")); print("cap $cap:", d["value"], d["ms_per_step"])
k=d["kernels"]
for n in ("pb_bn_act_bwd_reduce","pb_colstats","pb_bn_act_bwd_apply"): print("   %-24s %7.2f ms %7.0f GB/s" % (n, k[n]["ms_per_step"], k[n]["GBps"]))
PY
grep "bn_act_bwd_reduce" gpurun_out/detail_cap$cap.txt | sort -k1 -n -r | head -6
done

mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -rf -x > gpurun_out/pytest_r3m.txt 2>&1; tail -4 gpurun_out/pytest_r3m.txt | cut -c1-250
timeout 900 python -m pytest tests/test_blocks_bf16_gpu.py tests/test_models_gpu.py -m gpu -q --tb=short -rf -x 2>&1 | tail -2 | cut -c1-200
for v in on off; do
if [ $v = off ]; then export PB_GEMM_NO_PSRES=1; else unset PB_GEMM_NO_PSRES; fi
PB_BENCH_DETAIL=gpurun_out/detail_psres_$v.txt timeout 900 python bench.py --steps 4 --warmup 3 --no-torch-b200 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('psres $v:', d['value'], d['ms_per_step'], d['roofline']['frac'], k['pb_pw_gemm_tc']['ms_per_step'])"
done
grep "pb_pw_gemm_tc|64,64" gpurun_out/detail_psres_on.txt | head -8; echo; grep "pb_pw_gemm_tc|64,64" gpurun_out/detail_psres_off.txt | head -8

mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rf > gpurun_out/pytest_r2o.txt 2>&1; tail -8 gpurun_out/pytest_r2o.txt | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

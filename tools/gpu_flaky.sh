for i in 1 2 3; do timeout 1500 python -m pytest tests -m gpu -q --tb=line -rf 2>&1 | tail -2 | cut -c1-200; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-200

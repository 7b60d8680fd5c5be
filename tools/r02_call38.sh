set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python tools/sweep_dense_cherry.py 2>&1 | grep "^{" | cut -c1-200 | head -2

# verification of the round's last state: gpu tests, smoke, default bench
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r03q_bench_c2_1gpu.json 2> gpurun_out/r03q_bench_c2_1gpu.err; echo rc=$?

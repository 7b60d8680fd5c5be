set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/r03d_bench_c2_1gpu.json 2> gpurun_out/r03d_bench_c2_1gpu.err; echo rc=$?

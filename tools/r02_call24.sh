set -x
timeout 900 python -m pytest tests/test_inputs_gpu.py -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 > gpurun_out/r02v_bench_c2_1gpu.json 2> gpurun_out/r02v_bench_c2_1gpu.err; echo rc=$?

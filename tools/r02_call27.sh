set -x
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_masks_gpu.py tests/test_inputs_gpu.py -x -q -m gpu 2>&1 | tail -4
OPTS=fused_fold=0 python tools/ab_lib.py 2>&1 | tail -1 | cut -c1-420
OPTS=fused_fold=1 python tools/ab_lib.py 2>&1 | tail -1 | cut -c1-420
python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r02x_bench_c2_1gpu.json 2> gpurun_out/r02x_bench_c2_1gpu.err; echo rc=$?

set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "kernel_only or same_matrices or fused_dmma" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r03k_bench_c2_1gpu.json 2> gpurun_out/r03k_bench_c2_1gpu.err; echo rc=$?
tail -3 gpurun_out/r03k_bench_c2_1gpu.err | cut -c1-200

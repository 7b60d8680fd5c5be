set -x
python bench.py > gpurun_out/r02z_bench_c2_1gpu.json 2> gpurun_out/r02z_bench_c2_1gpu.err; echo rc=$?
tail -4 gpurun_out/r02z_bench_c2_1gpu.err | cut -c1-200

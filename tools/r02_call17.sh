# round 2, call 17: state check as the driver will see it: full gpu test tier, smoke, default bench (wall clock), reference arm
set -x
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) 2>&1
( time python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) 2>&1
( time python bench.py > gpurun_out/r02q_bench_c2_1gpu.json 2> gpurun_out/r02q_bench_c2_1gpu.err ) 2>&1; echo "rc=$?"
( time python bench.py --impl reference > gpurun_out/r02q_bench_reference_arm.json 2> gpurun_out/r02q_bench_reference_arm.err ) 2>&1; echo "rc=$?"
tail -3 gpurun_out/r02q_bench_c2_1gpu.err

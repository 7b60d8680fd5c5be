# final state of the round: tests, smoke, default bench, reference arm, launch list of the bench command
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r03h_bench_c2_1gpu.json 2> gpurun_out/r03h_bench_c2_1gpu.err; echo rc=$?
python bench.py --impl reference > gpurun_out/r03h_bench_reference_arm.json 2> gpurun_out/r03h_bench_reference_arm.err; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r03h_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r03h_ncu_launches.log 2>&1; echo rc=$?

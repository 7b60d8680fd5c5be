set -e
# launch list of the default bench, then one --set full capture of the top kernel and of the table builder (c2)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01n_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fused4c|build_clade' -s 10 -c 3 -o gpurun_out/r01n_fused4c_256x4_pow2_clades5 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1

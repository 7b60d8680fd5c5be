set -e
# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01l_launches_bench_default.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
# full capture of the dominant kernel (two consecutive category launches of fused4c after warm-up)
ncu --set full --clock-control none --import-source on -k regex:fused4c -s 8 -c 2 -o gpurun_out/r01l_fused4c_256x4_pow2_cherry -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1

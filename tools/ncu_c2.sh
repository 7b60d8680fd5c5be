set -e
# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01k_launches_bench_default.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
# full capture of the dominant kernel (one category launch of fused4c, 4th instance = after warm-up)
ncu --set full --clock-control none --import-source on -k regex:fused4c -s 8 -c 2 -o gpurun_out/r01k_fused4c_256x4_pow2 -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1

set -e
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01k_launches_bench_default.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused4c -s 8 -c 2 -o gpurun_out/r01k_fused4c_256x4_pow2 -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
for c in c3 c4; do
ncu --set full --clock-control none --import-source on -k regex:fused_dmma -s 3 -c 1 -o gpurun_out/r01k_${c}_fused_dmma3_pow2 -f python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/ncu_$c.log 2>&1
done

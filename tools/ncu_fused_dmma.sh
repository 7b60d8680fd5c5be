set -e
for c in c3 c4; do
ncu --set full --clock-control none --import-source on -k regex:fused_dmma -s 3 -c 1 -o gpurun_out/r01j_${c}_fused_dmma3 -f python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/ncu_$c.log 2>&1
done

# full captures of the final kernels, summarised on the box (the reports themselves are too large to bring back)
set -x
mkdir -p /tmp/cub && cd /tmp/cub && cuobjdump -xelf all $GRAFT_REPO_ROOT/phylogenetics_b200/libphylo_b200.so > /dev/null && nvdisasm -g -c pb_api.sm_100a.cubin > pb_api.txt 2>/dev/null; cd $GRAFT_REPO_ROOT
for c in c3 c4; do
ncu --set full --clock-control none --import-source on -k regex:fused_dmma3 -s 3 -c 1 -o /tmp/r03h_${c} -f python bench.py --config $c --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r03h_ncu_$c.log 2>&1; echo rc=$?
n=20; [ $c = c4 ] && n=61
{ echo "# ncu --set full --clock-control none, bench.py --config $c: fused_dmma3 (final code of round 2)"; python tools/ncu_summary.py /tmp/r03h_${c}.ncu-rep; echo; echo "# stall samples by CUDA source line"; python tools/ncu_lines.py /tmp/r03h_${c}.ncu-rep /tmp/cub/pb_api.txt fused_dmma3ILi${n}ELi2 150; } > gpurun_out/r03h_ncu_full_fused_dmma3_$c.txt 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:fused4t -s 2 -c 1 -o /tmp/r03h_fused4t -f python tools/run_fused_once.py 256 2 > gpurun_out/r03h_ncu_fused4t.log 2>&1; echo rc=$?
{ echo "# ncu --set full --clock-control none, tools/run_fused_once.py 256 2 (c2): fused4t, final code of round 2 (one kernel, FOLD body)"; python tools/ncu_summary.py /tmp/r03h_fused4t.ncu-rep; echo; echo "# stall samples by CUDA source line"; python tools/ncu_lines.py /tmp/r03h_fused4t.ncu-rep /tmp/cub/pb_api.txt fused4tILi256ELi2ELi0ELb0 150; } > gpurun_out/r03h_ncu_full_fused4t.txt 2>&1
ls -la gpurun_out | tail -12

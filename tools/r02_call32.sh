set -x
for o in "" "fused_clade_leaves=8" "fused_clade_leaves=6" "fused_spt=4" "fused_occ=4" "fused_block=128"; do OPTS=$o python tools/ab_lib.py 2>&1 | tail -1 | cut -c1-330; done | tee gpurun_out/r02z_ab_fold_options.jsonl
ncu --set full --clock-control none --import-source on -k regex:fused4t -s 4 -c 2 -o gpurun_out/r02z_fused4t_fold -f python tools/run_fused_once.py 256 2 > gpurun_out/r02z_ncu.log 2>&1
tail -2 gpurun_out/r02z_ncu.log
ls -la gpurun_out/*.ncu-rep

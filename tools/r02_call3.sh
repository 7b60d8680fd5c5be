# round 2, call 3: full GPU test suite + fused4t sweeps (difference coding, clade limit, shape)
set -x
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
for o in "" "fused_xor=0" "fused_clade_leaves=6" "fused_clade_leaves=7" "fused_clade_leaves=8" "fused_clade_leaves=6,fused_spt=2" "fused_clade_leaves=7,fused_spt=2" "fused_clade_leaves=6,fused_xor=0" "fused_clade_leaves=6,fused_const_cats=4"; do
  OPTS="$o" python tools/ab_lib.py 2>/dev/null | cut -c1-420
done | tee gpurun_out/r02c_ab_sweep.jsonl
export OPTS=fused_clade_leaves=6
python tools/run_fused_once.py 256 4 > gpurun_out/r02c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fused4t|build_clade' -s 10 -c 3 -o gpurun_out/r02c_fused4t_k6_xor -f python tools/run_fused_once.py 256 4 > gpurun_out/r02c_ncu.log 2>&1
tail -n 3 gpurun_out/r02c_plain.log gpurun_out/r02c_ncu.log

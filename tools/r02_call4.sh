# round 2, call 4: rest of the GPU suite, occupancy variants, the new bench line
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
for o in "fused_clade_leaves=7,fused_spt=2" "fused_clade_leaves=7,fused_spt=2,fused_occ=4" "fused_clade_leaves=6,fused_spt=2,fused_occ=4" "fused_clade_leaves=7,fused_spt=2,fused_occ=4,fused_const_cats=4" "fused_clade_leaves=7,fused_spt=2,fused_block=128"; do
  OPTS="$o" python tools/ab_lib.py 2>/dev/null | cut -c1-420
done | tee gpurun_out/r02d_ab_sweep.jsonl
python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
tail -n 12 gpurun_out/r02d_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02d_bench_ref.json 2> gpurun_out/r02d_bench_ref.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/r02d_bench_ref.json

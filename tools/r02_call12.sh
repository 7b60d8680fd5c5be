# round 2, call 12: fused root reduction A/B, full GPU suite, bench line
set -x
for o in "" "fused_root=0" "fused_clade_leaves=6" "fused_root=0,fused_const_cats=1"; do
  OPTS="$o" python tools/ab_lib.py 2>/dev/null | cut -c1-520
done | tee gpurun_out/r02l_ab_fused_root.jsonl
timeout 1800 python -m pytest tests -q -m gpu --maxfail=6 2>&1 | tail -12
python bench.py --steps 20 --warmup 3 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"
tail -n 6 gpurun_out/r02l_bench.err | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5

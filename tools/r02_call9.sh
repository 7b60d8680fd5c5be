# round 2, call 9: e2e option sweep, then the ncu evidence of the final kernels (launch list + full captures)
set -x
for o in "" "host_lead_cats=4,fused_const_cats=4" "host_lead_cats=4,fused_const_cats=2" "host_lead_cats=2,fused_const_cats=2" "host_lead_cats=3" "host_chunks=2,host_lead_cats=4,fused_const_cats=4" "host_chunk_align=0" "host_lead_cats=1"; do
  OPTS="$o" python tools/ab_lib.py 2>/dev/null | cut -c1-520
done | tee gpurun_out/r02i_ab_e2e_options.jsonl
# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02i_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_ncu_launches.log 2>&1
# full capture of the dominant kernel and of the table builder (c2, defaults: clades <= 7, 2 sites per thread)
python tools/run_fused_once.py 256 2 > gpurun_out/r02i_plain_fused.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fused4t|build_clade' -s 10 -c 3 -o gpurun_out/r02i_fused4t_256x2_clades7 -f python tools/run_fused_once.py 256 2 > gpurun_out/r02i_ncu_fused.log 2>&1
# dense alphabets with cherries tabulated
for c in c3 c4; do
python bench.py --config $c --steps 3 --warmup 3 --no-extras > gpurun_out/r02i_plain_$c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fused_dmma3|build_cherries' -s 2 -c 2 -o gpurun_out/r02i_${c}_fused_dmma3_cherries -f python bench.py --config $c --steps 3 --warmup 3 --no-extras > gpurun_out/r02i_ncu_$c.log 2>&1
done
# the level-scheduled kernel's DRAM traffic (three level launches)
python bench.py --path level --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_plain_level.log 2>&1 &&
ncu --set full --clock-control none -k regex:level4_async -s 60 -c 3 -o gpurun_out/r02i_level4_async -f python bench.py --path level --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_ncu_level.log 2>&1
tail -n 2 gpurun_out/r02i_plain_*.log gpurun_out/r02i_ncu_*.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -8

# round 2, call 10: ncu evidence of the final kernels, summarised ON THE BOX (the reports together exceed what gpurun copies back)
set -x
cd $GRAFT_REPO_ROOT
SO=phylogenetics_b200/libphylo_b200.so
mkdir -p /tmp/cub && (cd /tmp/cub && cuobjdump -xelf all $GRAFT_REPO_ROOT/$SO > /dev/null 2>&1; nvdisasm -g -c pb_api.sm_100a.cubin > /tmp/cub/pb_api.sass 2>/dev/null; ls -la /tmp/cub | head)
# 1. launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02j_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02j_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02j_ncu_launches.log 2>&1
# 2. the dominant kernel and the table builder (c2, defaults)
python tools/run_fused_once.py 256 2 > gpurun_out/r02j_plain_fused.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fused4t|build_clade' -s 3 -c 2 -o /tmp/r02j_fused4t -f python tools/run_fused_once.py 256 2 > gpurun_out/r02j_ncu_fused.log 2>&1
(echo "# ncu --set full --clock-control none, tools/run_fused_once.py 256 2 (c2: 1M sites, 128 taxa, GTR+G4; defaults: clades <= 7 leaves, 2 sites per thread, all 4 categories in one launch): build_clade_tables<true> and fused4t<256,2>"; python tools/ncu_summary.py /tmp/r02j_fused4t.ncu-rep; echo; echo "# stall samples by CUDA source line (tools/ncu_lines.py), the fused4t launch"; python tools/ncu_lines.py /tmp/r02j_fused4t.ncu-rep /tmp/cub/pb_api.sass fused4tILi256ELi2ELi0ELb0 120 fused4t) > gpurun_out/r02j_ncu_full_fused4t_256x2_clades7.txt 2>&1
ncu -i /tmp/r02j_fused4t.ncu-rep --page raw --csv > gpurun_out/r02j_fused4t_raw.csv 2>/dev/null
# 3. dense alphabets with cherries tabulated
for c in c3 c4; do
python bench.py --config $c --steps 3 --warmup 3 --no-extras > gpurun_out/r02j_plain_$c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fused_dmma3|build_cherries' -s 2 -c 2 -o /tmp/r02j_$c -f python bench.py --config $c --steps 3 --warmup 3 --no-extras > gpurun_out/r02j_ncu_$c.log 2>&1
(echo "# ncu --set full --clock-control none, bench.py --config $c: fdmma3_build_cherries and fused_dmma3 (cherries tabulated)"; python tools/ncu_summary.py /tmp/r02j_$c.ncu-rep; echo; echo "# stall samples by CUDA source line, fused_dmma3"; python tools/ncu_lines.py /tmp/r02j_$c.ncu-rep /tmp/cub/pb_api.sass fused_dmma3ILi$([ $c = c3 ] && echo 20 || echo 61)ELi2 100 fused_dmma3) > gpurun_out/r02j_ncu_full_fused_dmma3_cherries_$c.txt 2>&1
done
# 4. the level-scheduled kernel's DRAM traffic (three level launches)
python bench.py --path level --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02j_plain_level.log 2>&1 &&
ncu --set full --clock-control none -k regex:level4_async -s 60 -c 3 -o /tmp/r02j_level -f python bench.py --path level --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02j_ncu_level.log 2>&1
(echo "# ncu --set full --clock-control none, bench.py --path level: three launches of level4_async"; python tools/ncu_summary.py /tmp/r02j_level.ncu-rep) > gpurun_out/r02j_ncu_full_level4_async.txt 2>&1
tail -n 2 gpurun_out/r02j_plain_*.log | cut -c1-300
du -sh gpurun_out

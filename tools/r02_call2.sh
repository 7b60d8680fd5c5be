# round 2, call 2: fused4t (reduced program) parity + A/B against fused4c on c2
set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fused_reduced or fused_constant or pipelined_host or two_contexts or fused_variants" 2>&1 | tail -15
python tools/ab_lib.py > gpurun_out/r02b_ab_fused4t.json 2> gpurun_out/r02b_ab_fused4t.err
OPTS=fused_reduced=0 python tools/ab_lib.py > gpurun_out/r02b_ab_fused4c.json 2> gpurun_out/r02b_ab_fused4c.err
OPTS=fused_clade_leaves=6 python tools/ab_lib.py > gpurun_out/r02b_ab_fused4t_k6.json 2>/dev/null
OPTS=fused_clade_leaves=4 python tools/ab_lib.py > gpurun_out/r02b_ab_fused4t_k4.json 2>/dev/null
OPTS=fused_const_cats=4 python tools/ab_lib.py > gpurun_out/r02b_ab_fused4t_cats4.json 2>/dev/null
OPTS=fused_spt=2 python tools/ab_lib.py > gpurun_out/r02b_ab_fused4t_spt2.json 2>/dev/null
cat gpurun_out/r02b_ab_*.json | cut -c1-330
tail -5 gpurun_out/r02b_ab_fused4t.err
python tools/run_fused_once.py 256 4 > gpurun_out/r02b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused4t -s 8 -c 2 -o gpurun_out/r02b_fused4t -f python tools/run_fused_once.py 256 4 > gpurun_out/r02b_ncu.log 2>&1
tail -n 3 gpurun_out/r02b_plain.log gpurun_out/r02b_ncu.log

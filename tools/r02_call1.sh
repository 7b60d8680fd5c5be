# round 2, call 1: FUSED_PREAPPLY A/B on c2 + ncu full capture (source-level) of the preapply kernel
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python tools/ab_lib.py > gpurun_out/r02a_ab_default.json 2> gpurun_out/r02a_ab_default.err
PHYLO_B200_LIB=build/libphylo_pre.so python tools/ab_lib.py > gpurun_out/r02a_ab_prelib_off.json 2> gpurun_out/r02a_ab_prelib_off.err
PHYLO_B200_LIB=build/libphylo_pre.so OPTS=fused_preapply=1 python tools/ab_lib.py > gpurun_out/r02a_ab_preapply.json 2> gpurun_out/r02a_ab_preapply.err
cat gpurun_out/r02a_ab_*.json
export PHYLO_B200_LIB=build/libphylo_pre.so OPTS=fused_preapply=1
python tools/run_fused_once.py 256 4 > gpurun_out/r02a_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused4c -s 8 -c 2 -o gpurun_out/r02a_fused4c_preapply -f python tools/run_fused_once.py 256 4 > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_plain.log gpurun_out/r02a_ncu.log

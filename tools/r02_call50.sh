set -x
mkdir -p /tmp/cub && cd /tmp/cub && cuobjdump -xelf all $GRAFT_REPO_ROOT/phylogenetics_b200/libphylo_b200.so > /dev/null && nvdisasm -g -c pb_api.sm_100a.cubin > pb_api.txt 2>/dev/null; cd $GRAFT_REPO_ROOT
python tools/run_masked_once.py 2>&1 | tail -1
OPTS=fused_clade_leaves=6 python tools/run_fused_once.py 256 2 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:fused4t -s 2 -c 1 -o /tmp/r03j_masked -f python tools/run_masked_once.py > gpurun_out/r03j_ncu_masked.log 2>&1; echo rc=$?
{ echo "# ncu --set full --clock-control none, tools/run_masked_once.py (c2, 10 % of the cells missing): fused4t<MASKS>, FOLD body"; python tools/ncu_summary.py /tmp/r03j_masked.ncu-rep; echo; echo "# stall samples by CUDA source line"; python tools/ncu_lines.py /tmp/r03j_masked.ncu-rep /tmp/cub/pb_api.txt fused4tILi256ELi2ELi0ELb1 150; } 2>&1 | grep -v "^+ " > gpurun_out/r03j_ncu_full_fused4t_masks.txt

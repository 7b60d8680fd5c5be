set -x
mkdir -p /tmp/cub && cd /tmp/cub && cuobjdump -xelf all $GRAFT_REPO_ROOT/phylogenetics_b200/libphylo_b200.so > /dev/null && nvdisasm -g -c pb_api.sm_100a.cubin > pb_api.txt 2>/dev/null; cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:fdmma3_build_cherries -s 3 -c 1 -o /tmp/r03n_builder -f python bench.py --config c4 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r03n_ncu.log 2>&1; echo rc=$?
{ echo "# ncu --set full, bench.py --config c4: fdmma3_build_cherries<61> (one wave of blocks)"; python tools/ncu_summary.py /tmp/r03n_builder.ncu-rep; echo; python tools/ncu_lines.py /tmp/r03n_builder.ncu-rep /tmp/cub/pb_api.txt fdmma3_build_cherriesILi61 30; } 2>&1 | grep -v "^+ " > gpurun_out/r03n_ncu_full_build_cherries_c4.txt

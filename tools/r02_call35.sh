set -x
python tools/taxa_probe.py 1024 "" "fused_fold=0" "fused_clade_leaves=6" "fused_spt=1" 2>&1 | grep "^{" | tee gpurun_out/r02z_taxa_probe_1024.jsonl
python tools/taxa_probe.py 512 "" "fused_fold=0" 2>&1 | grep "^{" | tee -a gpurun_out/r02z_taxa_probe_1024.jsonl

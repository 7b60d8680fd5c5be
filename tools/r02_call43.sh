set -x
python tools/sweep_dense_opts.py c3 "" "fused_dmma_slots=3" "fused_dmma_slots=4" "fused_dmma_slots=5" "fused_dmma_slots=6" "fused_dmma_slots=7" 2>&1 | grep "^{" | tee gpurun_out/r03e_sweep_c3_slots.jsonl

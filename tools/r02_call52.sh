set -x
python tools/sweep_dense_opts.py c4 "" "fused_dmma_builder_waves=1" "fused_dmma_builder_waves=2" "fused_dmma_builder_waves=3" "fused_dmma_builder_waves=4" 2>&1 | grep "^{" | tee gpurun_out/r03l_sweep_builder_waves.jsonl
python tools/sweep_dense_opts.py c3 "" "fused_dmma_builder_waves=1" "fused_dmma_builder_waves=2" 2>&1 | grep "^{" | tee -a gpurun_out/r03l_sweep_builder_waves.jsonl
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fused_dmma or dense" 2>&1 | tail -2

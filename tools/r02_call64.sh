set -x
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_reference_golden_gpu.py tests/test_inputs_gpu.py -x -q -m gpu -k "dmma or dense or golden" 2>&1 | tail -3
python tools/sweep_dense_opts.py c4 "" "fused_dmma_builder_waves=2" 2>&1 | grep "^{"
python tools/sweep_dense_opts.py c3 "" 2>&1 | grep "^{"

set -x
timeout 900 python -m pytest tests/test_inputs_gpu.py tests/test_parity_gpu.py -x -q -m gpu -k "schedule or fused_dmma" 2>&1 | tail -3

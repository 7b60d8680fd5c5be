set -x
timeout 900 python -m pytest tests/test_masks_gpu.py tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/run_masked_once.py 2>&1 | tail -1
python tools/run_fused_once.py 256 2 2>&1 | tail -1

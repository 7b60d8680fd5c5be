set -x
python tools/sweep_dense_cherry.py 2>&1 | grep "^{" | cut -c1-260 | tee gpurun_out/r03c_sweep_dense_2ctas.jsonl
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_reference_golden_gpu.py tests/test_reference_api_gpu.py -q -m gpu -x 2>&1 | tail -3

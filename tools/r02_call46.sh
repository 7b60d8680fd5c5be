set -x
python tools/e2e_dense_probe.py "" "host_chunks=1" 2>&1 | grep "^{" | tee gpurun_out/r03g_e2e_dense_rampup.jsonl | cut -c1-330
timeout 900 python -m pytest tests/test_inputs_gpu.py tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -3

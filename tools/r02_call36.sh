set -x
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_masks_gpu.py tests/test_inputs_gpu.py tests/test_sharded_abi_gpu.py -x -q -m gpu 2>&1 | tail -4
for o in "" "fused_fold=0"; do OPTS=$o python tools/ab_lib.py 2>&1 | tail -1 | cut -c1-330; done
python tools/e2e_trace.py "" 2>&1 | cut -c1-160

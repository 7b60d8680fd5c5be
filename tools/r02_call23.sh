set -x
timeout 900 python -m pytest tests/test_inputs_gpu.py tests/test_parity_gpu.py tests/test_sharded_abi_gpu.py -x -q -m gpu 2>&1 | tail -4
python tools/e2e_trace.py "" "host_side_streams=0" "host_chunks=6" "host_chunks=8" "host_chunks=12" "host_chunks=16" 2>&1 | tee gpurun_out/r02u_e2e_trace.jsonl | cut -c1-200

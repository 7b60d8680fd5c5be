set -x
timeout 900 python -m pytest tests/test_inputs_gpu.py tests/test_parity_gpu.py tests/test_sharded_abi_gpu.py tests/test_masks_gpu.py -x -q -m gpu 2>&1 | tail -4
python tools/e2e_trace.py "" "host_side_streams=0" "host_chunks=6" "host_chunks=8" "host_chunks=6,host_tail_sites=-1" "host_chunks=3" "host_chunks=6,host_tail_sites=120000" 2>&1 | tee gpurun_out/r02s_e2e_trace.jsonl | cut -c1-200

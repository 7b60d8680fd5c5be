set -x
timeout 900 python -m pytest tests/test_inputs_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/e2e_trace.py "" "host_chunks=8" "host_chunks=4" "host_tail_sites=-1" 2>&1 | tee gpurun_out/r02z_e2e_trace.jsonl | cut -c1-140

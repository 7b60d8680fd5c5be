set -x
python tools/e2e_trace.py "" "host_copy_streams=2" "host_copy_streams=2,host_chunks=8" "" "host_copy_streams=2" 2>&1 | tee gpurun_out/r03b_e2e_trace_copy_streams.jsonl | cut -c1-160

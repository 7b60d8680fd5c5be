set -x
python tools/e2e_trace.py "" "host_lanes=3" "host_lanes=4" "host_lanes=3,host_chunks=8" "host_lanes=3,host_chunks=12" "" "host_lanes=3" 2>&1 | tee gpurun_out/r03p_e2e_trace_lanes.jsonl | cut -c1-140

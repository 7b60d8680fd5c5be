set -x
python tools/e2e_trace.py "" "host_chunks=8" "host_chunks=5" "host_tail_sites=114000" 2>&1 | tee gpurun_out/r03o_e2e_trace_final.jsonl | cut -c1-120

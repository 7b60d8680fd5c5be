# round 2, call 18: timeline of the pipelined host evaluation under several chunk schedules
set -x
python tools/e2e_trace.py "" "host_chunks=8" "host_chunks=16" "host_chunks=4,host_tail_sites=57000" "host_chunks=4,host_tail_sites=114000" "host_chunks=8,host_tail_sites=57000" "host_chunks=2,host_tail_sites=57000" "host_chunks=4,host_chunk_align=0" 2>&1 | tee gpurun_out/r02r_e2e_trace.jsonl | cut -c1-1500

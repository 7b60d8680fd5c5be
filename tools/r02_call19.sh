set -x
python tools/split_probe.py 2>&1 | tee gpurun_out/r02r_split_probe.jsonl
python tools/e2e_trace.py "" "fused_const_keep=0" 2>&1 | tee gpurun_out/r02r_e2e_trace2.jsonl | cut -c1-300

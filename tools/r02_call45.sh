set -x
python tools/e2e_dense_probe.py "host_chunks=1" "host_chunks=2" "host_chunks=3" "host_chunks=4" "host_chunks=6" "host_chunks=2,host_tail_sites=-1" "host_chunks=3,host_tail_sites=-1" 2>&1 | grep "^{" | tee gpurun_out/r03f_e2e_dense_probe.jsonl | cut -c1-330

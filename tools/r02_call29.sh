set -x
python -c "
import sys; sys.path.insert(0,'.')
from phylogenetics_b200.engine import device_local_cpus
import os
print('local cpus', sorted(device_local_cpus(0) or [])[:4], len(device_local_cpus(0) or []), 'affinity', len(os.sched_getaffinity(0)))
"
nvidia-smi topo -m 2>&1 | head -14
for i in 1 2 3; do PB_E2E_TRACE=1 python bench.py --steps 20 --warmup 3 --no-extras 2> gpurun_out/r02y_bench_$i.err | tee gpurun_out/r02y_bench_$i.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['reps_ms_per_step'], d['e2e']['byte_per_site']['ms_per_step'])"; done
python tools/e2e_trace.py "" "host_chunks=8" 2>&1 | tee gpurun_out/r02y_e2e_trace.jsonl | cut -c1-240

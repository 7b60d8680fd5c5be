set -x
timeout 600 python -m pytest tests/test_transition_matrices_gpu.py -x -q -m gpu 2>&1 | tail -3
python -c "
import sys; sys.path.insert(0,'.')
from phylogenetics_b200.engine import measure_fp64_peak
for _ in range(3): print(measure_fp64_peak(0))
"
python bench.py --no-extras --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['fp64_peaks_live'])"

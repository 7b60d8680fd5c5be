set -x
timeout 900 python -m pytest tests/test_inputs_gpu.py -x -q -m gpu 2>&1 | tail -12
python bench.py --config c3 --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['ms_per_step'], d['e2e'])"
python bench.py --config c4 --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4', d['ms_per_step'], d['e2e'])"

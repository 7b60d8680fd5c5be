set -x
timeout 1500 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -4
for i in 1 2; do PB_E2E_TRACE=1 python bench.py --steps 20 --warmup 3 --no-extras 2> gpurun_out/r02x_bench_$i.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['byte_per_site']['ms_per_step'])"; grep timeline gpurun_out/r02x_bench_$i.err | cut -c1-900; done

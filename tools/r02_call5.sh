# round 2, call 5: masks in fused4t, sharded ABI, remaining GPU tests, bench line with new defaults
set -x
nvidia-smi -L
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
tail -n 8 gpurun_out/r02e_bench.err

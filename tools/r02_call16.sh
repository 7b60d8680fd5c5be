# round 2, call 16: three-leaf tables at 20 states, dense parity tests, bench blocks
set -x
python tools/sweep_dense_cherry.py 2>&1 | tee gpurun_out/r02p_sweep_dense_cherry_triple.jsonl | cut -c1-260
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_reference_golden_gpu.py tests/test_reference_api_gpu.py -q -m gpu -k "dmma or dense or parity or golden or reference" --maxfail=4 2>&1 | tail -5
python bench.py --config c3 --steps 10 --warmup 3 --no-extras > gpurun_out/r02p_bench_c3.json 2> gpurun_out/r02p_bench_c3.err; echo "rc=$?"

# round 2, call 13: cherry builder with staged matrices, fused4t after the fused-root revert, full suite, bench
set -x
python tools/sweep_dense_cherry.py 2>&1 | tee gpurun_out/r02m_sweep_dense_cherry.jsonl | cut -c1-260
OPTS="" python tools/ab_lib.py 2>/dev/null | cut -c1-520 | tee gpurun_out/r02m_ab_default.jsonl
timeout 1800 python -m pytest tests -q -m gpu --maxfail=6 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo "bench rc=$?"
tail -n 6 gpurun_out/r02m_bench.err | cut -c1-300

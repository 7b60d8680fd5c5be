# round 2, call 15: cherry builder (256 threads, staged matrices), dense parity tests, c3 / c4 bench blocks
set -x
python tools/sweep_dense_cherry.py 2>&1 | tee gpurun_out/r02o_sweep_dense_cherry.jsonl | cut -c1-260
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "dmma or dense or parity" --maxfail=4 2>&1 | tail -5
python bench.py --config c4 --steps 10 --warmup 3 --no-extras > gpurun_out/r02o_bench_c4.json 2> gpurun_out/r02o_bench_c4.err; echo "rc=$?"
python bench.py --config c3 --steps 10 --warmup 3 --no-extras > gpurun_out/r02o_bench_c3.json 2> gpurun_out/r02o_bench_c3.err; echo "rc=$?"
python bench.py --config c4 --steps 3 --warmup 3 --no-extras > gpurun_out/r02o_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fused_dmma3|build_cherries' -s 2 -c 2 -o /tmp/r02o_c4 -f python bench.py --config c4 --steps 3 --warmup 3 --no-extras > gpurun_out/r02o_ncu_c4.log 2>&1
(echo "# ncu --set full --clock-control none, bench.py --config c4: fdmma3_build_cherries<61> (256 threads, matrices staged in shared memory) and fused_dmma3<61,2>"; python tools/ncu_summary.py /tmp/r02o_c4.ncu-rep) 2>&1 | grep -v "^+" > gpurun_out/r02o_ncu_full_fused_dmma3_cherries_c4.txt
grep -E "Kernel Name|time_duration" gpurun_out/r02o_ncu_full_fused_dmma3_cherries_c4.txt | cut -c1-140

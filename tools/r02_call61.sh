# two GPUs: the sharded C-ABI tests that skip on one device, and the 2-rank bench
set -x
timeout 600 python -m pytest tests/test_sharded_abi_gpu.py -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r03r_bench_2gpu.json 2> gpurun_out/r03r_bench_2gpu.err; echo rc=$?
tail -2 gpurun_out/r03r_bench_2gpu.err | cut -c1-200

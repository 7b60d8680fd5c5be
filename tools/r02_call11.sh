# round 2, call 11 (8 GPUs): the scaling line with the single-process arm and the config-5 block, concurrent H2D probe
set -x
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02k_bench_8gpu.json 2> gpurun_out/r02k_bench_8gpu.err; echo "bench rc=$?"
tail -n 4 gpurun_out/r02k_bench_8gpu.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/h2d_probe_multi.py > gpurun_out/r02k_h2d_probe_8ranks.json 2>/dev/null; cat gpurun_out/r02k_h2d_probe_8ranks.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 3 --no-extras > gpurun_out/r02k_bench_4gpu.json 2> gpurun_out/r02k_bench_4gpu.err; echo "bench4 rc=$?"
timeout 600 python -m pytest tests/test_sharded_abi_gpu.py -q -m gpu 2>&1 | tail -3

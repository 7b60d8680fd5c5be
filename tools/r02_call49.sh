# 8 GPUs of one box, final code: the scaling bench as the driver launches it
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r03i_bench_8gpu.json 2> gpurun_out/r03i_bench_8gpu.err; echo rc=$?
tail -2 gpurun_out/r03i_bench_8gpu.err | cut -c1-300

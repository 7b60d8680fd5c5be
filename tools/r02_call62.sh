# 4 GPUs of one box, final code
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r03s_bench_4gpu.json 2> gpurun_out/r03s_bench_4gpu.err; echo rc=$?
tail -1 gpurun_out/r03s_bench_4gpu.err | cut -c1-200

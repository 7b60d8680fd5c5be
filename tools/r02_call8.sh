# round 2, call 8 (2 GPUs): multi-GPU inside the C ABI, torchrun bench with the single-process arm and the config-5 block
set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_sharded_abi_gpu.py tests/test_sharded_optimise_gpu.py tests/test_ocaml_stubs.py "tests/test_parity_gpu.py::test_fused_reduced_program_tree_sizes" -q -m gpu 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.err; echo "bench rc=$?"
tail -n 6 gpurun_out/r02h_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/h2d_probe_multi.py > gpurun_out/r02h_h2d_probe_2ranks.json 2>/dev/null; cat gpurun_out/r02h_h2d_probe_2ranks.json
timeout 300 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-200

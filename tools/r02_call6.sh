# round 2, call 6: full GPU suite (masks, sharded ABI, graph, OCaml stubs, dense cherries), e2e chunk sweep, c5 loop on one GPU
set -x
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python tools/sweep_dense_cherry.py 2>&1 | tee gpurun_out/r02f_sweep_dense_cherry.jsonl | cut -c1-260
LEADS=1,2,4 python tools/sweep_e2e_packed.py 2>/dev/null | tee gpurun_out/r02f_sweep_e2e_chunks_x_lead_cats.jsonl | cut -c1-140
python tools/optimise_c5.py --sites-per-gpu 1250000 --evaluations 400 > gpurun_out/r02f_c5_1gpu_graph.json 2> gpurun_out/r02f_c5.err; echo "c5 rc=$?"
python tools/optimise_c5.py --sites-per-gpu 1250000 --evaluations 400 --no-graph > gpurun_out/r02f_c5_1gpu_nograph.json 2>> gpurun_out/r02f_c5.err
cut -c1-700 gpurun_out/r02f_c5_1gpu_graph.json gpurun_out/r02f_c5_1gpu_nograph.json; tail -n 5 gpurun_out/r02f_c5.err

# round 2, call 7: the OCaml stubs driver's full output, the GPU suite (all failures), the bench line
set -x
python -c "
import sys; sys.path.insert(0,'tests')
import test_ocaml_stubs as t; print(t.build_driver())"
tests/ocaml_mock/_build/stubs_driver > gpurun_out/r02g_stubs_driver.txt 2>&1; echo "driver rc=$?"
grep -n "FAIL" gpurun_out/r02g_stubs_driver.txt
timeout 1800 python -m pytest tests -q -m gpu --maxfail=6 2>&1 | tail -25
python bench.py --steps 20 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"
tail -n 8 gpurun_out/r02g_bench.err

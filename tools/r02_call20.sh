set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02r_launches_e2e_trace.csv python tools/e2e_trace.py "" > gpurun_out/r02r_ncu_e2e.log 2>&1
tail -2 gpurun_out/r02r_ncu_e2e.log | cut -c1-200

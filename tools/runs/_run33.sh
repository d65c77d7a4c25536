mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 520 --csv --log-file gpurun_out/r02_launches_v3.csv python tools/one_step.py > gpurun_out/ncu_ll3.log 2>&1; tail -1 gpurun_out/ncu_ll3.log; wc -l gpurun_out/r02_launches_v3.csv

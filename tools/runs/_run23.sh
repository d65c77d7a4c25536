mkdir -p gpurun_out
echo "== step profile"; timeout 200 python tools/step_profile.py tc 2>&1 | grep -v Warn | head -30 | tee gpurun_out/r02_step_profile_v2.txt
echo "== launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "step2/" --csv --log-file gpurun_out/r02_launches_v2.csv python tools/one_step.py > gpurun_out/ncu_ll.log 2>&1; tail -1 gpurun_out/ncu_ll.log; wc -l gpurun_out/r02_launches_v2.csv
echo "== ncu split kernels"; timeout 400 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "step2/" -k regex:plane_split -c 40 -o gpurun_out/r02_split python tools/one_step.py > gpurun_out/ncu_split.log 2>&1; tail -1 gpurun_out/ncu_split.log

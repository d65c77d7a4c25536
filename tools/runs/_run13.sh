mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_audio.py tests/test_gpu_train.py -q -x 2>&1 | tail -12; [ ${PIPESTATUS[0]} -ne 0 ] && { echo STOP audio/train; }
echo "== preset models strict"; timeout 900 python -m pytest tests/test_gpu_models.py -q -s -k preset 2>&1 | grep -E "worst|passed|failed|Error|assert " | head -30
echo "== incremental + dropin + rest"; timeout 900 python -m pytest tests/test_gpu_incremental.py tests/test_dropin.py tests/test_gpu_models.py -q -k "not preset" 2>&1 | tail -4
echo "== launch list of one eager step (ncu, cold-cache serialised)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_tc_eager.csv python tools/one_step.py > gpurun_out/one_step.log 2>&1; tail -2 gpurun_out/one_step.log; wc -l gpurun_out/r02_launches_tc_eager.csv

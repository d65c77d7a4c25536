mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_modules_golden.py tests/test_gpu_fusion.py tests/test_gpu_train.py -q -x 2>&1 | tail -6; [ ${PIPESTATUS[0]} -ne 0 ] && { echo STOP basic; exit 1; }
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== dropin"; timeout 400 python -m pytest tests/test_dropin.py -q -x -m gpu 2>&1 | tail -25
echo "== preset models strict"; timeout 900 python -m pytest tests/test_gpu_models.py -q -s -k preset 2>&1 | grep -E "worst|passed|failed|Error|error|assert" | head -40
echo "== bench default"; timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== step profile"; timeout 200 python tools/step_profile.py tc 2>&1 | head -26

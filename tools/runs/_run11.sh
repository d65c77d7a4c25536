mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_fusion.py tests/test_gpu_modules_golden.py tests/test_gpu_train.py -q -x 2>&1 | tail -8; [ ${PIPESTATUS[0]} -ne 0 ] && { echo STOP basic; exit 1; }
echo "== preset models strict"; timeout 900 python -m pytest tests/test_gpu_models.py -q -s -k preset 2>&1 | grep -E "worst|passed|failed|Error|assert " | head -30
echo "== attention standalone"; timeout 120 python tools/attn_time.py
echo "== bench no-extras: attention A/B"; for a in 1 0; do DV3_TC_ATTN=$a timeout 200 python bench.py --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('TC_ATTN=$a', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"; done
echo "== ncu attention"; ATTN_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 3 -c 3 -o gpurun_out/r02_ncu_attn python tools/attn_time.py > gpurun_out/ncu_attn.log 2>&1; tail -3 gpurun_out/ncu_attn.log
echo "== ncu conv"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 2 -c 2 -o gpurun_out/r02_ncu_conv python tools/profile_convblock.py > gpurun_out/ncu_conv.log 2>&1; tail -3 gpurun_out/ncu_conv.log

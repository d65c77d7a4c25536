mkdir -p gpurun_out
echo "== fuse debug swz0"; timeout 60 python tools/fuse_debug.py 2>&1 | tail -12
timeout 150 python -m pytest tests/test_gpu_fusion.py -q -x 2>&1 | tail -15; if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "fusion failed; trying swizzle variants"; for s in 1 2 3; do echo "== swz $s"; DV3_DEBUG_SWZ=$s timeout 60 python tools/fuse_debug.py 2>&1 | tail -8; done; exit 1; fi
timeout 500 python -m pytest tests/test_gpu_modules_golden.py tests/test_gpu_train.py tests/test_gpu_models.py -q -x 2>&1 | tail -15; [ ${PIPESTATUS[0]} -ne 0 ] && { echo STOP models; exit 1; }
for gm in 0.56; do echo "== trunc bias gamma=$gm"; DV3_TC_GAMMA=$gm timeout 100 python tools/trunc_bias.py 2>&1 | grep "tc "; done
for gm in 0 0.56; do echo "== precision gamma=$gm"; DV3_TC_GAMMA=$gm timeout 300 python tools/precision_presets.py --math tc --no64 2>&1 | grep -E "==|gpu_tc"; done
echo "== bench default"; timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== bench FUSE off"; DV3_FUSE_FWD=0 DV3_FUSE_BWD=0 timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== bench FUSE bwd off"; DV3_FUSE_BWD=0 timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== step profile"; timeout 200 python tools/step_profile.py tc 2>&1 | head -22

mkdir -p gpurun_out
echo "== gpu tests: blocks fusion modules train"
timeout 1500 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_fusion.py tests/test_gpu_modules_golden.py tests/test_gpu_train.py -q -x 2>&1 | tail -30
for gm in 0 0.46; do echo "== trunc bias gamma=$gm"; DV3_TC_GAMMA=$gm timeout 200 python tools/trunc_bias.py 2>&1 | grep "tc "; done
for gm in 0 0.46; do echo "== precision gamma=$gm"; DV3_TC_GAMMA=$gm timeout 400 python tools/precision_presets.py --math tc --no64 2>&1 | grep -E "==|gpu_tc"; done
echo "== bench default"; timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== bench FUSE off"; DV3_FUSE_FWD=0 DV3_FUSE_BWD=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== bench FUSE bwd off"; DV3_FUSE_BWD=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== step profile"; timeout 300 python tools/step_profile.py tc 2>&1 | head -22

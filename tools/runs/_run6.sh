mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@"; rc=$?; if [ $rc -ne 0 ]; then echo "FAILED rc=$rc -- stopping"; exit 1; fi; }
run "blocks" 240 python -m pytest tests/test_gpu_blocks.py -q -x 2>&1 | tail -15
[ ${PIPESTATUS[0]} -ne 0 ] && exit 1
timeout 120 python -m pytest tests/test_gpu_fusion.py -q -x 2>&1 | tail -15; [ ${PIPESTATUS[0]} -ne 0 ] && { echo STOP fusion; exit 1; }
timeout 400 python -m pytest tests/test_gpu_modules_golden.py tests/test_gpu_train.py tests/test_gpu_models.py -q -x 2>&1 | tail -15; [ ${PIPESTATUS[0]} -ne 0 ] && { echo STOP models; exit 1; }
for gm in 0.56 ; do echo "== trunc bias gamma=$gm"; DV3_TC_GAMMA=$gm timeout 100 python tools/trunc_bias.py 2>&1 | grep "tc "; done
for gm in 0 0.56; do echo "== precision gamma=$gm"; DV3_TC_GAMMA=$gm timeout 300 python tools/precision_presets.py --math tc --no64 2>&1 | grep -E "==|gpu_tc"; done
echo "== bench default"; timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== bench FUSE off"; DV3_FUSE_FWD=0 DV3_FUSE_BWD=0 timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== bench FUSE bwd off"; DV3_FUSE_BWD=0 timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
echo "== step profile"; timeout 200 python tools/step_profile.py tc 2>&1 | head -22

mkdir -p gpurun_out
echo "== loss test"; timeout 300 python -m pytest tests/test_gpu_train.py -q 2>&1 | tail -3
for gm in 0 1.0; do echo "== trunc bias gamma=$gm"; DV3_TC_GAMMA=$gm timeout 200 python tools/trunc_bias.py 2>&1 | tail -8; done
for gm in 0.7 1.0 1.4; do echo "== precision gamma=$gm"; DV3_TC_GAMMA=$gm timeout 400 python tools/precision_presets.py deepvoice3_ljspeech deepvoice3_vctk --math tc --no64 2>&1 | grep -E "==|gpu_tc"; done
echo "== step profile fuse on";  timeout 300 python tools/step_profile.py tc 2>&1 | head -24
echo "== step profile fuse off"; DV3_FUSE_FWD=0 DV3_FUSE_BWD=0 timeout 300 python tools/step_profile.py tc 2>&1 | head -24
echo "== bench FWD only"; DV3_FUSE_BWD=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "== bench BWD only"; DV3_FUSE_FWD=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200

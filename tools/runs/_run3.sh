mkdir -p gpurun_out
echo "== gpu tests (PDL on): modules golden, train, blocks, fusion"
timeout 1200 python -m pytest tests/test_gpu_modules_golden.py tests/test_gpu_train.py tests/test_gpu_blocks.py tests/test_gpu_fusion.py -q 2>&1 | tail -40
echo "== precision"
timeout 600 python tools/precision_presets.py > gpurun_out/r02_precision0.txt 2>&1
echo "precision rc=$?"
tail -20 gpurun_out/r02_precision0.txt
echo "== bench default"; timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
echo "== bench PDL=0"; DV3_PDL=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
echo "== bench FUSE off"; DV3_FUSE_FWD=0 DV3_FUSE_BWD=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400

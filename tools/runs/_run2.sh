mkdir -p gpurun_out
echo "== attention / module / train golden tests (PDL on)"
timeout 900 python -m pytest tests/test_gpu_modules_golden.py tests/test_gpu_train.py tests/test_gpu_blocks.py -x -q 2>&1 | tail -15
echo "== same, DV3_PDL=0"
DV3_PDL=0 timeout 900 python -m pytest tests/test_gpu_modules_golden.py tests/test_gpu_train.py -x -q 2>&1 | tail -8
echo "== precision"
timeout 600 python tools/precision_presets.py > gpurun_out/r02_precision0.txt 2>&1
echo "precision rc=$?"
tail -20 gpurun_out/r02_precision0.txt
echo "== pair64"
timeout 900 bash tools/pair64_check.sh > gpurun_out/r02_pair64.txt 2>&1
cat gpurun_out/r02_pair64.txt
echo "== bench PDL=0"; DV3_PDL=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300

mkdir -p gpurun_out
timeout 600 python tools/precision_presets.py > gpurun_out/r02_precision0.txt 2>&1
echo "precision rc=$?"
timeout 900 bash tools/pair64_check.sh > gpurun_out/r02_pair64.txt 2>&1
echo "pair64 rc=$?"
tail -20 gpurun_out/r02_precision0.txt
cat gpurun_out/r02_pair64.txt

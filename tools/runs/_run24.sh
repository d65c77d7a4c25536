mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_modules_golden.py tests/test_gpu_blocks.py -x -q -k "attn or attention" 2>&1 | tail -3
for v in A B A B; do echo "lib$v"; DV3_LIB=$PWD/tools/ab/lib$v.so ATTN_REPS=20 timeout 120 python tools/attn_time.py 2>&1 | grep "^tc"; done | tee gpurun_out/r02_attn_ab.log

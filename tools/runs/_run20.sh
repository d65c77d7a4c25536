mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_audio.py -x -q 2>&1 | tail -3
timeout 120 python tools/stft_time.py 2>&1 | tail -2 | tee gpurun_out/r02_stft_time.log

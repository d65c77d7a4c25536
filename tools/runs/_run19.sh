mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_audio.py -x -q 2>&1 | tail -15
timeout 120 python tools/stft_time.py 2>&1 | tail -2 | tee gpurun_out/r02_stft_time.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_mel -c 1 -s 3 -o gpurun_out/r02_stft_v5 python tools/stft_time.py > gpurun_out/ncu_stft.log 2>&1; tail -2 gpurun_out/ncu_stft.log

mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r02_gputests.log
timeout 120 python tools/stft_time.py 2>&1 | tail -2 | tee gpurun_out/r02_stft_time.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_mel -c 1 -s 3 -o gpurun_out/r02_stft python tools/stft_time.py > gpurun_out/ncu_stft.log 2>&1; tail -2 gpurun_out/ncu_stft.log

mkdir -p gpurun_out
timeout 200 python tools/stft_ab.py A B 2>&1 | tail -4 | tee gpurun_out/r02_stft_ab.log

mkdir -p gpurun_out
timeout 200 python tools/stft_ab.py A B 2>&1 | tail -4 | tee gpurun_out/r02_stft_ab2.log
DV3_LIB=$PWD/tools/ab/libB.so timeout 300 python -m pytest tests/test_gpu_audio.py -x -q 2>&1 | tail -2

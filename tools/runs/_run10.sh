mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_fusion.py -q -x 2>&1 | tail -8; [ ${PIPESTATUS[0]} -ne 0 ] && { echo STOP basic; exit 1; }
echo "== preset models strict"; timeout 900 python -m pytest tests/test_gpu_models.py -q -s -k preset 2>&1 | grep -E "worst|passed|failed|Error|assert " | head -30
echo "== bench no-extras: attention A/B"; for a in 1 0; do DV3_TC_ATTN=$a timeout 200 python bench.py --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('TC_ATTN=$a', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"; done
echo "== bench PDL off"; DV3_PDL=0 timeout 200 python bench.py --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL=0', d['ms_per_step'])"
echo "== full bench"; timeout 900 python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; echo rc=$?; tail -c 3000 gpurun_out/r02_bench_full.json; tail -5 gpurun_out/r02_bench_full.err
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400

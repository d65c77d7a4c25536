mkdir -p gpurun_out
( timeout 500 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -4 )
for v in 0 1 0 1; do
  DV3_OVERLAP_WN=$v timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 40 --warmup 5 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('OVERLAP_WN=$v ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'launches', d.get('gpu_launches'))"
done 2>&1 | tee gpurun_out/r02_overlap_wn_ab.log

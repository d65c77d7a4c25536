mkdir -p gpurun_out
for v in 0 6 0 6 12; do
  DV3_TC_PREFETCH=$v timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 40 --warmup 5 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PREFETCH=$v ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), [ (s['C'],s['T'],round(s['fwd_us'],1),round(s['dgrad_us'],1)) for s in d['roofline']['shapes']])"
done 2>&1 | tee gpurun_out/r02_prefetch_ab.log

mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r02_gputests_final.log
echo "== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench full"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_full_v2.json 2> gpurun_out/r02_bench_n1_full_v2.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_n1_full_v2.json

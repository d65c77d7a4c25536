mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | tee gpurun_out/r02_gputests_final.log
echo "== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1

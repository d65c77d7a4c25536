mkdir -p gpurun_out
echo "== pytest -m gpu (whole suite)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -4
echo "== precision"; timeout 300 python tools/precision_presets.py --math tc --no64 2>&1 | grep -E "==|gpu_tc"
echo "== step profile"; timeout 200 python tools/step_profile.py tc 2>&1 | grep -v Warn | head -14
echo "== bench default (no extras)"; timeout 200 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"

timeout 400 python -m pytest tests/test_dropin.py -x -q -k "synthesis" 2>&1 | tail -15

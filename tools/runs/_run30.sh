timeout 400 python -m pytest tests/test_dropin.py -x -q -k "ljspeech_preprocessor" 2>&1 | tail -15

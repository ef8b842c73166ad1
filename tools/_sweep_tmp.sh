timeout 300 python -m pytest -m gpu -x -q tests/test_resample.py 2>&1 | tail -5
timeout 300 python tools/time_resample.py 2>&1 | tail -2

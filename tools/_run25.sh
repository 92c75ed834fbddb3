timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "many_short" 2>&1 | tail -30

timeout 400 python -m pytest tests/test_gpu_display.py tests/test_gpu_parity.py -x -q 2>&1 | tail -12

timeout 300 python -m pytest tests/test_gpu_display.py tests/test_host_traces.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -5

timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_display.py tests/test_gpu_sharded.py tests/test_gpu_device.py -x -q -k "minmax or fake or decimate or special or compressed or wholefile or ops_match" 2>&1 | tail -3
timeout 100 python tools/kbench.py --op minmax --C 8 --step 1920 --steps 30
timeout 100 python tools/kbench.py --op minmax --C 4 --rate 96000 --seconds 160 --step 1382400 --steps 30
timeout 100 python tools/kbench.py --op minmax --C 64 --rate 250000 --seconds 4 --step 500 --steps 30
timeout 100 python tools/kbench.py --op minmax --C 1 --rate 44100 --seconds 600 --step 4410 --steps 30
timeout 100 python tools/kbench.py --op minmax --C 1 --rate 44100 --seconds 600 --step 4411 --steps 30
timeout 100 python tools/kbench.py --op minmax --C 16 --rate 500000 --seconds 8 --step 150000 --steps 30
timeout 100 python tools/kbench.py --op minmax --C 2 --rate 48000 --seconds 320 --step 8000 --steps 30

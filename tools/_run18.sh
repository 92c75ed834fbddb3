timeout 400 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -x -q -k "envelope or golden" 2>&1 | tail -5
timeout 100 python tools/kbench.py --op envelope --C 8 --steps 20
timeout 100 python tools/kbench.py --op envelope --C 64 --rate 250000 --seconds 4 --steps 10
timeout 100 python tools/kbench.py --op envelope --C 1 --rate 44100 --seconds 600 --steps 10

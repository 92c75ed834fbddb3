timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device.py -x -q -k "not spectrogram" 2>&1 | tail -3
timeout 100 python tools/kbench.py --op filter --C 8 --order 4 --steps 20
timeout 100 python tools/kbench.py --op filter --C 8 --steps 20
timeout 100 python tools/kbench.py --op filter --C 8 --order 2 --kind lowpass --steps 20
timeout 100 python tools/kbench.py --op envelope --C 8 --steps 20
timeout 100 python tools/kbench.py --op filter --C 64 --rate 250000 --seconds 4 --order 4
timeout 100 python tools/kbench.py --op filter --C 64 --rate 250000 --seconds 4

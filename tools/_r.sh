timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k spectrogram 2>&1 | tail -3
timeout 100 python tools/kbench.py --op spectrogram --steps 20
timeout 100 python tools/kbench.py --op spectrogram --steps 20 --hop 256
timeout 100 python tools/kbench.py --op spectrogram --C 64 --rate 250000 --seconds 4 --steps 10

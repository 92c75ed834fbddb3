timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 100 python tools/kbench.py --op filter --C 64 --rate 250000 --seconds 4
timeout 100 python tools/kbench.py --op filter --C 64 --rate 250000 --seconds 4 --order 4
timeout 100 python tools/kbench.py --op envelope --C 64 --rate 250000 --seconds 4
timeout 100 python tools/kbench.py --op filter --C 8 --rate 48000 --seconds 80
timeout 100 python tools/kbench.py --op filter --C 8 --rate 48000 --seconds 80 --order 4
timeout 100 python tools/kbench.py --op filter --C 8 --rate 48000 --seconds 80 --order 2 --kind lowpass
timeout 100 python tools/kbench.py --op envelope --C 8 --rate 48000 --seconds 80
timeout 100 python tools/kbench.py --op filter --C 4 --rate 96000 --seconds 160 --order 4

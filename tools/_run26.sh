timeout 100 python tools/kbench.py --op envelope --C 8 --steps 30
timeout 100 python tools/kbench.py --op envelope --C 8 --steps 30
timeout 100 python tools/kbench.py --op filter --C 8 --order 2 --kind lowpass --steps 30

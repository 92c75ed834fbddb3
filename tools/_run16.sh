timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k spectrogram 2>&1 | tail -6
for n in 2048 4096 8192 16384; do
timeout 100 python tools/kbench.py --op spectrogram --C 8 --nfft $n --hop $((n/2)) --steps 10
timeout 100 python tools/kbench.py --op spectrogram --C 64 --rate 250000 --seconds 4 --nfft $n --hop $((n/2)) --steps 5
done

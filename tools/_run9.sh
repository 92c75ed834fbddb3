timeout 100 python tools/kbench.py --op spectrogram --C 16 --rate 500000 --seconds 8
timeout 100 python tools/kbench.py --op spectrogram --C 2 --rate 48000 --seconds 320
timeout 100 python tools/kbench.py --op spectrogram --C 1 --rate 44100 --seconds 600
timeout 100 python tools/kbench.py --op minmax --C 4 --rate 96000 --seconds 160 --step 1382400
timeout 100 python tools/kbench.py --op filter --C 4 --rate 96000 --seconds 160 --order 4
timeout 400 python tools/sweep.py --seconds 4 --out gpurun_out/sweep_c5.json

ncu --set full --clock-control none --import-source on -k regex:spectrogram_mp -s 3 -c 1 -o gpurun_out/prof_mp -f python tools/kbench.py --op spectrogram --nfft 2048 --hop 1024 --steps 3 > gpurun_out/ncu_mp.log 2>&1
ncu -i gpurun_out/prof_mp.ncu-rep --page source --csv > gpurun_out/src_mp.csv 2>/dev/null
ncu -i gpurun_out/prof_mp.ncu-rep --page raw --csv > gpurun_out/raw_mp.csv 2>/dev/null

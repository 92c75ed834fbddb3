export ADN_SPEC_NW=4 ADN_SPEC_CB=4
ncu --set full --clock-control none --import-source on -k regex:spectrogram_ring -s 3 -c 1 -o gpurun_out/prof_spec_ring -f python tools/kbench.py --op spectrogram --steps 3 > gpurun_out/ncu_ring.log 2>&1
ncu -i gpurun_out/prof_spec_ring.ncu-rep --page source --csv > gpurun_out/src_ring.csv 2>/dev/null
ncu -i gpurun_out/prof_spec_ring.ncu-rep --page raw --csv > gpurun_out/raw_ring.csv 2>/dev/null

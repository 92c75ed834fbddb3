ncu --set full --clock-control none --import-source on -k regex:minmax_split -s 3 -c 1 -o gpurun_out/prof_mm -f python tools/kbench.py --op minmax --step 1920 --steps 3 > gpurun_out/ncu_mm.log 2>&1
ncu -i gpurun_out/prof_mm.ncu-rep --page source --csv > gpurun_out/src_mm.csv 2>/dev/null
ncu -i gpurun_out/prof_mm.ncu-rep --page raw --csv > gpurun_out/raw_mm.csv 2>/dev/null

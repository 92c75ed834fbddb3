ncu --set full --clock-control none --import-source on -k regex:sos_scan -s 3 -c 1 -o gpurun_out/prof_s2 -f python tools/kbench.py --op filter --steps 3 > gpurun_out/ncu_s2.log 2>&1
ncu -i gpurun_out/prof_s2.ncu-rep --page source --csv > gpurun_out/src_s2.csv 2>/dev/null
ncu -i gpurun_out/prof_s2.ncu-rep --page raw --csv > gpurun_out/raw_s2.csv 2>/dev/null

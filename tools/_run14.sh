ncu --set full --clock-control none --import-source on -k regex:sos_scan -s 3 -c 1 -o gpurun_out/prof_s4 -f python tools/kbench.py --op filter --order 4 --steps 3 > gpurun_out/ncu_s4.log 2>&1
ncu -i gpurun_out/prof_s4.ncu-rep --page source --csv > gpurun_out/src_s4.csv 2>/dev/null
ncu -i gpurun_out/prof_s4.ncu-rep --page raw --csv > gpurun_out/raw_s4.csv 2>/dev/null

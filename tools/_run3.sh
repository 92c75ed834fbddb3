python -m pytest tests/test_gpu_parity.py -x -q -k spectrogram 2>&1 | tail -5
for cfg in "4 4" "4 2" "4 8" "2 2"; do
  set -- $cfg
  echo "NW=$1 CB=$2"
  ADN_SPEC_NW=$1 ADN_SPEC_CB=$2 python tools/kbench.py --op spectrogram --steps 20
done
export ADN_SPEC_NW=4 ADN_SPEC_CB=4
ncu --set full --clock-control none --import-source on -k regex:spectrogram_ring -s 3 -c 1 -o gpurun_out/prof_spec_ring -f python tools/kbench.py --op spectrogram --steps 3 > gpurun_out/ncu_ring.log 2>&1
ncu -i gpurun_out/prof_spec_ring.ncu-rep --page source --csv > gpurun_out/src_ring.csv 2>/dev/null
ncu -i gpurun_out/prof_spec_ring.ncu-rep --page raw --csv > gpurun_out/raw_ring.csv 2>/dev/null

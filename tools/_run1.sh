set -x
python -m pytest tests/test_gpu_parity.py -x -q -k spectrogram 2>&1 | tail -15
for cfg in "1 8 8" "1 4 4" "1 8 4" "1 4 8" "0 8 8"; do
  set -- $cfg
  echo "RING=$1 NW=$2 CB=$3"
  ADN_SPEC_RING=$1 ADN_SPEC_NW=$2 ADN_SPEC_CB=$3 python tools/kbench.py --op spectrogram --steps 20
done

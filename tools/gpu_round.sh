#!/bin/bash
# Standard GPU pass of a development round (run through gpurun):
#   tests, bench, ncu launch list of the bench, one full capture per hot kernel.
# Usage: bash tools/gpu_round.sh <tag>      -> gpurun_out/<tag>_*
tag=${1:-r01}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> $out/${tag}_tests.log
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_ref.json 2>> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 > $out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'spectrogram_ring|sos_scan' -s 12 -c 4 \
    -o $out/${tag}_full -f python bench.py --steps 2 --warmup 3 > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_full.ncu-rep --page raw --csv > $out/${tag}_full_raw.csv 2>/dev/null

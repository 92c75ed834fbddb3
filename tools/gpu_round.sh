#!/bin/bash
# Standard GPU pass of a development round (run through gpurun):
#   tests, bench (ours + reference arm), ncu launch list of the bench, one full capture per hot kernel.
# Usage: bash tools/gpu_round.sh <tag>      -> gpurun_out/<tag>_*
tag=${1:-r02}
out=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_ref.json 2>> $out/${tag}_bench.err
# the launch list is restricted to the timed region of the bench (NVTX range "timed")
timeout 900 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-wholefile > $out/${tag}_ncu_launch.log 2>&1
timeout 1200 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none --import-source on -k regex:'spectrogram_ring|sos_scan|sos_run|sos_zp|sos_fwd' -c 3 \
    -o $out/${tag}_full -f python bench.py --steps 2 --warmup 3 --no-wholefile > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_full.ncu-rep --page raw --csv > $out/${tag}_full_raw.csv 2>/dev/null
ncu -i $out/${tag}_full.ncu-rep --page source --csv -k regex:spectrogram_ring > $out/${tag}_ring_src.csv 2>/dev/null
for o in 1 2 3 4; do timeout 300 python tools/kbench.py --op filter --order $o --steps 30 2>/dev/null | tail -1; done > $out/${tag}_kbench.jsonl
timeout 300 python tools/kbench.py --op filter --order 4 --C 4 --rate 96000 --seconds 160 --steps 20 2>/dev/null | tail -1 >> $out/${tag}_kbench.jsonl
timeout 300 python tools/kbench.py --op envelope --steps 30 2>/dev/null | tail -1 >> $out/${tag}_kbench.jsonl
timeout 300 python tools/kbench.py --op spectrogram --steps 30 2>/dev/null | tail -1 >> $out/${tag}_kbench.jsonl
timeout 300 python tools/kbench.py --op minmax --step 1920 --steps 30 2>/dev/null | tail -1 >> $out/${tag}_kbench.jsonl
timeout 900 python tools/sweep.py --out $out/${tag}_sweep_c5.json > /dev/null 2>&1
echo done

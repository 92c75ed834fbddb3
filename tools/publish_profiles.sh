#!/bin/bash
# Copy the evidence of one tools/gpu_round.sh pass from gpurun_out/ (scratch) into profiles/ (tracked).
#   tools/publish_profiles.sh <tag in gpurun_out> <name in profiles>
set -e
tag=$1; name=$2
cd "$(dirname "$0")/.."
g=gpurun_out; p=profiles
tail -1 $g/${tag}_bench.json > $p/${name}_bench.json
tail -1 $g/${tag}_ref.json > $p/${name}_ref.json
grep -v '^==' $g/${tag}_launches.csv > $p/${name}_launches.csv
python profiles/summarize.py launches $g/${tag}_launches.csv > $p/${name}_launches.txt
python tools/ncu_digest.py raw $g/${tag}_full_raw.csv > $p/${name}_full.txt
python tools/opmix.py $g/${tag}_ring_src.csv 60000 frame > $p/${name}_ring_opmix.txt
cp $g/${tag}_kbench.jsonl $p/${name}_kbench.jsonl
cp $g/${tag}_sweep_c5.json $p/${name}_sweep_c5.json
tail -3 $g/${tag}_tests.log > $p/${name}_tests.txt
ls -la $p/${name}_*

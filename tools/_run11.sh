bash tools/gpu_round.sh r01d
timeout 300 python tools/sweep.py --seconds 4 --out gpurun_out/r01d_sweep_c5.json > gpurun_out/r01d_sweep.log 2>&1
timeout 200 python tools/wholefile_bench.py --config 4 --seconds 7200 > gpurun_out/r01d_wholefile.jsonl 2>&1
timeout 200 python tools/wholefile_bench.py --config 3 --seconds 240 >> gpurun_out/r01d_wholefile.jsonl 2>&1
for op in "filter --C 8" "envelope --C 8" "spectrogram --C 8" "minmax --C 8 --step 1920" "filter --C 8 --order 4" "spectrogram --C 1 --rate 44100 --seconds 600" "spectrogram --C 16 --rate 500000 --seconds 8" "minmax --C 4 --rate 96000 --seconds 160 --step 1382400" "filter --C 4 --rate 96000 --seconds 160 --order 4"; do
timeout 100 python tools/kbench.py --op $op --steps 20 >> gpurun_out/r01d_kbench.jsonl 2>&1
done

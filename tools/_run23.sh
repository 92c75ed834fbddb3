python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r01e_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r01e_smoke.log
bash tools/gpu_round.sh r01e
timeout 300 python tools/sweep.py --seconds 4 --out gpurun_out/r01e_sweep_c5.json > gpurun_out/r01e_sweep.log 2>&1
timeout 200 python tools/wholefile_bench.py --config 4 --seconds 7200 > gpurun_out/r01e_wholefile.jsonl 2>&1
timeout 200 python tools/wholefile_bench.py --config 3 --seconds 240 >> gpurun_out/r01e_wholefile.jsonl 2>&1
rm -f gpurun_out/r01e_kbench.jsonl
for op in "filter --C 8" "filter --C 8 --order 2 --kind lowpass" "filter --C 8 --order 4" "envelope --C 8" "spectrogram --C 8" "minmax --C 8 --step 1920" "spectrogram --C 1 --rate 44100 --seconds 600" "spectrogram --C 16 --rate 500000 --seconds 8" "spectrogram --C 8 --nfft 4096 --hop 2048" "minmax --C 4 --rate 96000 --seconds 160 --step 1382400" "filter --C 4 --rate 96000 --seconds 160 --order 4" "filter --C 64 --rate 250000 --seconds 4"; do
timeout 100 python tools/kbench.py --op $op --steps 20 >> gpurun_out/r01e_kbench.jsonl 2>&1
done

set -x
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_sharded.py -x -q -k "filt or envelope or sos or shard" > gpurun_out/a_tests.log 2>&1; echo "rc=$?" >> gpurun_out/a_tests.log
for o in "--op filter" "--op filter --order 1 --kind lowpass" "--op filter --order 4" "--op envelope" "--op filter --C 64 --rate 250000 --seconds 4" "--op filter --C 4 --rate 96000 --seconds 160 --order 4"; do
  timeout 120 python tools/kbench.py $o --steps 20
done > gpurun_out/a_kbench.jsonl 2>&1
timeout 200 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err

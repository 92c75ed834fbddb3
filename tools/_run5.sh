python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/kbench.py --op envelope --steps 20
python tools/kbench.py --op filter --steps 20
python bench.py --steps 20 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -3 gpurun_out/bench5.err

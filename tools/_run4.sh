python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -3 gpurun_out/bench4.err

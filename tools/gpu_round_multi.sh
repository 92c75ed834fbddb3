#!/bin/bash
# Multi-GPU pass (run through gpurun --gpus N): bench.py under torchrun at N ranks + the NCCL test.
#   bash tools/gpu_round_multi.sh <N>   -> gpurun_out/b2_bench_n<N>.json
out=gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $out/b2_bench_n$N.json 2> $out/b2_bench_n$N.err; echo "bench rc=$?"
tail -c 800 $out/b2_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/b2_bench_n$N.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','op_ms','parity','minmax','e2e','gpu_launches','dominant'):
        print(k, json.dumps(d.get(k))[:600])
    for k,v in (d.get('wholefile') or {}).items(): print(k, json.dumps(v)[:1200])
except Exception as e: print('parse failed', e)
PY
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k nccl 2>&1 | tail -3

timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "rc=$?"; tail -2 gpurun_out/bench_n4.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 tools/wholefile_bench.py --gpus 4 --config 4 --seconds 28800 > gpurun_out/wholefile_n4.jsonl 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29546 tools/wholefile_bench.py --gpus 4 --config 3 --seconds 900 >> gpurun_out/wholefile_n4.jsonl 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29547 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/ref_n4.json 2>&1; echo "ref rc=$?"

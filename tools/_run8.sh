timeout 200 python tools/wholefile_bench.py --config 4 --seconds 3600
timeout 200 python tools/wholefile_bench.py --config 3 --seconds 120

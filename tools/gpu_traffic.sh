#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_s7.json 2> gpurun_out/bench_s7.err
python -c "import json; l=json.loads(open('gpurun_out/bench_s7.json').read().strip().splitlines()[-1]); print('BENCH', l['value'], l['ms_per_step'], l['e2e'])"
tail -3 gpurun_out/bench_s7.err
timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/traffic_s7.csv python tools/profile_step.py 64 hrnet 1 > gpurun_out/ncu_s7.log 2>&1
tail -2 gpurun_out/ncu_s7.log

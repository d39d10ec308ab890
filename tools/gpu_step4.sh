#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L="c48,c96,c192,c384,b1x1b,b1x1a,d96,f1x1"
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_hrnet.py tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-250
timeout 300 python tools/conv_layer_bench.py 64 1 $L 2>&1 | grep "conv_test" | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH', l['value'], l['ms_per_step'], l['roofline']['ms'], l['e2e']['value'])"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_s4.csv python tools/profile_step.py 64 hrnet 1 > gpurun_out/ncu_s4.log 2>&1
tail -2 gpurun_out/ncu_s4.log

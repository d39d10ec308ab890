#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SHAPY_CONV_DEBUG=1 timeout 300 python tools/conv_layer_bench.py 64 1 c48,c96,c192,c384 2>&1 | grep "conv_test\|halo\]\|per-tap\]" | cut -c1-200
echo "== NT=64 on c192 / NT=128,64 on c384"
SHAPY_CONV_NT=64 timeout 300 python tools/conv_layer_bench.py 64 1 c192,c384 2>&1 | grep "conv_test" | cut -c1-200
SHAPY_CONV_NT=128 timeout 300 python tools/conv_layer_bench.py 64 1 c384 2>&1 | grep "conv_test" | cut -c1-200
echo "== rowsched=1"
SHAPY_CONV_ROWSCHED=1 timeout 300 python tools/conv_layer_bench.py 64 1 c48,c96 2>&1 | grep "conv_test" | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH', l['value'], l['ms_per_step'], l['roofline']['ms'], l['e2e']['value'])"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_s2.csv python tools/profile_step.py 64 hrnet 1 > gpurun_out/ncu_s2.log 2>&1
tail -2 gpurun_out/ncu_s2.log

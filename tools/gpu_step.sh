#!/bin/bash
# one optimisation step: conv + hrnet parity, per-layer timings, bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_hrnet.py tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-250
SHAPY_CONV_DEBUG=1 timeout 300 python tools/conv_layer_bench.py 64 1 $1 2>&1 | grep "conv_test\|halo\]" | cut -c1-260
timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH', l['value'], l['ms_per_step'], l['roofline']['ms'], l['e2e']['value'])"

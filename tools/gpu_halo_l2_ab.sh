#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_hrnet.py -x -q -m gpu 2>&1 | tail -2 | cut -c1-250
L="c48,c96,d48,d96,d192,s2_96_192"
echo "== l2_bpc 40 (default)"
SHAPY_CONV_DEBUG=1 timeout 200 python tools/conv_layer_bench.py 64 1 $L 2>&1 | grep "conv_test\|halo\] cin" | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH', l['value'], l['ms_per_step'], l['roofline']['ms'], l['e2e']['value'])"
echo "== l2_bpc 19.6"
SHAPY_HALO_L2BPC=19.6 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH', l['value'], l['ms_per_step'], l['roofline']['ms'], l['e2e']['value'])"

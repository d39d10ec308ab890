#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c15; mkdir -p $O
timeout 180 python -m pytest tests/test_gpu_smplx.py -q -m gpu -x 2>&1 | tail -6 | cut -c1-300
echo "--- TV off"; SHAPY_LBS_TV=0 timeout 120 python tools/lbs_time.py 64 2>&1 | tail -1
echo "--- TV on"; timeout 120 python tools/lbs_time.py 64 256 4096 2>&1 | tail -3
SHAPY_LBS_DEBUG=1 timeout 120 python tools/lbs_time.py 64 2>&1 | grep "lbs\]" | tail -15
timeout 300 python -m pytest tests/test_gpu_e2e.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300

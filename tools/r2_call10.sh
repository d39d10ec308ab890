#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c10; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_smplx.py tests/test_gpu_e2e.py -q -m gpu -x 2>&1 | tail -4
SHAPY_LBS_DEBUG=1 timeout 120 python tools/lbs_time.py 64 2>&1 | grep "lbs\]" | tail -16
timeout 120 python tools/lbs_time.py 64 256 4096 2>&1 | tail -4
SHAPY_LBS_FUSED=0 timeout 120 python tools/lbs_time.py 64 4096 2>&1 | tail -3

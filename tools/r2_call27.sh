#!/bin/bash
cd "$(dirname "$0")/.."
timeout 90 python tools/lbs_time.py 64 2>&1 | tail -1 || { echo "HUNG/FAILED quick run"; exit 1; }
timeout 300 python -m pytest tests/test_gpu_zz_edges.py tests/test_gpu_smplx.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300
timeout 120 python tools/lbs_time.py 64 256 4096 2>&1 | tail -3
SHAPY_LBS_DEBUG=1 timeout 60 python tools/lbs_time.py 64 2>&1 | grep "lbs\]" | tail -20

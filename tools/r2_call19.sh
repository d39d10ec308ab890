#!/bin/bash
cd "$(dirname "$0")/.."
timeout 90 python tools/lbs_time.py 64 2>&1 | tail -1 || { echo "HUNG/FAILED quick run"; exit 1; }
timeout 180 python -m pytest tests/test_gpu_smplx.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300
echo "--- TV on"; timeout 120 python tools/lbs_time.py 256 4096 2>&1 | tail -2
SHAPY_LBS_DEBUG=1 timeout 60 python tools/lbs_time.py 64 2>&1 | grep "lbs\]" | tail -15
timeout 200 python -m pytest tests/test_gpu_e2e.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300

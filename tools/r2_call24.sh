#!/bin/bash
cd "$(dirname "$0")/.."
timeout 90 python tools/lbs_time.py 64 2>&1 | tail -1 || { echo "HUNG/FAILED quick run"; exit 1; }
timeout 180 python -m pytest tests/test_gpu_smplx.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300
echo "--- TV on"; LBS_PROF=1 timeout 120 python tools/lbs_time.py 64 256 4096 2>&1 | grep -v "^$" | tail -9
echo "--- TV off"; SHAPY_LBS_TV=0 timeout 120 python tools/lbs_time.py 64 4096 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_e2e.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300

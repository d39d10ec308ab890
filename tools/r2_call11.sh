#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c11; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_smplx.py tests/test_gpu_e2e.py -q -m gpu -x 2>&1 | tail -4
SHAPY_LBS_DEBUG=1 timeout 120 python tools/lbs_time.py 64 2>&1 | grep "lbs\]" | tail -15
timeout 120 python tools/lbs_time.py 64 256 4096 2>&1 | tail -4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"smplx_lbs|smplx_joints" -c 4 -o $O/lbs_full -f python tools/profile_step.py 64 lbs 1 > $O/ncu_lbs.log 2>&1; echo "ncu rc $?"

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c22; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:smplx_lbs_kernel --launch-skip 3 --launch-count 1 -f -o $O/lbs_tv_B4096 python tools/lbs_time.py 4096 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ls -la $O

#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- TV on"; LBS_PROF=1 timeout 120 python tools/lbs_time.py 64 4096 2>&1 | tail -12
echo "--- clean flush"; LBS_CLEAN_FLUSH=1 timeout 120 python tools/lbs_time.py 64 4096 2>&1 | tail -2
SHAPY_LBS_DEBUG=1 timeout 120 python tools/lbs_time.py 64 2>&1 | grep "lbs\]" | tail -32

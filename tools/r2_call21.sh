#!/bin/bash
cd "$(dirname "$0")/.."
for f in 0 1 2 3; do echo "--- flags $f"; SHAPY_LBS_DBGFLAGS=$f timeout 90 python tools/lbs_time.py 4096 2>&1 | tail -1; done
SHAPY_LBS_DBGFLAGS=1 SHAPY_LBS_DEBUG=1 timeout 90 python tools/lbs_time.py 4096 2>&1 | grep "lbs\]" | grep "it0\|it1\|exit" | tail -14

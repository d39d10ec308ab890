#!/bin/bash
cd "$(dirname "$0")/.."
SHAPY_LBS_DEBUG=1 timeout 90 python tools/lbs_time.py 4096 2>&1 | grep "lbs\]" | tail -24

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c6; mkdir -p $O
timeout 120 python tools/measure_diag.py 2>&1 | tail -4
run() { n=$1; shift; env "$@" timeout 300 python tools/e2e_repro.py 64 10 u8 > $O/repro_$n.log 2>&1; echo "$n rc $? $(grep -c 'ok' $O/repro_$n.log) $(grep -m1 -i 'error' $O/repro_$n.log | cut -c1-160)"; }
run default X=1
run nograph SHAPY_HRNET_GRAPH=0
run nolbs SHAPY_LBS_FUSED=0
run measv1 SHAPY_MEASURE_V1=1
run lanes1 SHAPY_HRNET_LANES=1 SHAPY_HRNET_GRAPH=0
env X=1 timeout 300 python tools/e2e_repro.py 64 10 f32 > $O/repro_f32.log 2>&1; echo "f32 rc $? $(grep -c ok $O/repro_f32.log)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/e2e_repro.py 4 6 u8 > $O/sanitizer.log 2>&1; echo "sanitizer rc $?"; grep -E "Invalid|at 0x|in |ERROR SUMMARY|by thread" $O/sanitizer.log | head -30

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $1 FORCE=$2"; SHAPY_HALO_FORCE=$2 SHAPY_CONV_DEBUG=1 timeout 120 python tools/conv_layer_bench.py 64 1 $1 2>&1 | grep "conv_test\|halo\] cin" | cut -c1-230; }
echo "== defaults"; SHAPY_CONV_DEBUG=1 timeout 120 python tools/conv_layer_bench.py 64 1 c48,c96,d48,d96 2>&1 | grep "conv_test\|halo\] cin" | cut -c1-230
run c48 48,1,2
run c48 48,1,6
run c48 48,1,8
run c96 96,1,4
run c96 48,1,7
run c96 96,1,14
run c96 32,1,14
run d48 48,1,7
run d48 48,1,14
run d96 96,1,14
run d96 48,1,14

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
L="d96,d48,d192,s2_96_192"
echo "== halo s2"
SHAPY_CONV_DEBUG=1 timeout 300 python tools/conv_layer_bench.py 64 1 $L 2>&1 | grep "conv_test\|halo\]" | cut -c1-220
echo "== per-tap s2"
SHAPY_CONV_HALO_S2_MAXKCH=0 timeout 300 python tools/conv_layer_bench.py 64 1 $L 2>&1 | grep "conv_test" | cut -c1-200
echo "== halo s2 kch64"
SHAPY_CONV_HALO_S2_MAXKCH=64 SHAPY_CONV_DEBUG=1 timeout 300 python tools/conv_layer_bench.py 64 1 s2_64,s2_192_384,s2_256_96 2>&1 | grep "conv_test\|halo\]" | cut -c1-220
SHAPY_CONV_HALO_S2_MAXKCH=0 timeout 300 python tools/conv_layer_bench.py 64 1 s2_64,s2_192_384,s2_256_96 2>&1 | grep "conv_test" | cut -c1-200

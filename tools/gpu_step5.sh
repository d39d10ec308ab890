#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q -m gpu 2>&1 | tail -2 | cut -c1-250
echo "== KPAD=1 parity"
SHAPY_CONV_KPAD=1 timeout 900 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "split" 2>&1 | tail -3 | cut -c1-250
L="c48,d96,d48,d192,f1x1,s2_96_192"
echo "== default"
timeout 300 python tools/conv_layer_bench.py 64 1 $L 2>&1 | grep "conv_test" | cut -c1-200
echo "== KPAD=1"
SHAPY_CONV_KPAD=1 timeout 300 python tools/conv_layer_bench.py 64 1 $L 2>&1 | grep "conv_test" | cut -c1-200

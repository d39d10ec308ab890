#!/bin/bash
# bring-up of the TMA / swizzled-row A slices of the halo kernel: parity with and without the base-offset field,
# with the 64-channel layers forced through the halo kernel too, then per-layer timings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for bo in 1 0; do
  echo "=== parity BASEOFF=$bo (halo for kch<=32)"
  SHAPY_CONV_BASEOFF=$bo timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "0-shape0 or 0-shape1 or 0-shape12 or fp16" 2>&1 | tail -4 | cut -c1-200
  echo "=== parity BASEOFF=$bo MAXKCH=64"
  SHAPY_CONV_BASEOFF=$bo SHAPY_CONV_HALO_MAXKCH=64 timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "0-shape2 or 0-shape3 or 0-shape5" 2>&1 | tail -4 | cut -c1-200
done
echo "=== timings TMA"
timeout 300 python tools/conv_layer_bench.py 64 1 c48,c96 2>&1 | grep conv_test
SHAPY_CONV_TEST_REPS=1 SHAPY_CONV_PHASES=1 timeout 300 python tools/conv_layer_bench.py 64 1 c48,c96 2>&1 | grep phases | awk "NR%2==0"
echo "=== timings TMA MAXKCH=64"
SHAPY_CONV_HALO_MAXKCH=64 timeout 300 python tools/conv_layer_bench.py 64 1 c192,c384,b64,t48 2>&1 | grep conv_test
echo "=== timings legacy"
SHAPY_CONV_HALO_TMA=0 timeout 300 python tools/conv_layer_bench.py 64 1 c48,c96 2>&1 | grep conv_test

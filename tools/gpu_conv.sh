#!/bin/bash
# tcgen05 conv bring-up: SIMT engine first, then each tcgen05 shape in its own process (hang => timeout).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== SIMT engine"
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "1-shape or unsupported" > gpurun_out/conv_simt.log 2>&1
echo "exit $?"; tail -n 5 gpurun_out/conv_simt.log
for i in 0 1 2 3 4 5 6 7 8 9 10 11 12; do
  echo "=== tcgen05 shape$i"
  timeout 120 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "0-shape$i and split" > gpurun_out/conv_umma_$i.log 2>&1
  rc=$?
  echo "exit $rc"; tail -n 4 gpurun_out/conv_umma_$i.log | cut -c1-300
  if [ $rc -eq 124 ]; then echo "TIMEOUT (hang) on shape$i - stopping"; nvidia-smi > gpurun_out/smi_after_hang.txt 2>&1; break; fi
done
echo "=== fp16 mode"
timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "fp16" > gpurun_out/conv_fp16.log 2>&1
echo "exit $?"; tail -n 4 gpurun_out/conv_fp16.log | cut -c1-300

#!/bin/bash
# Runs the GPU test files one process per file (a faulting kernel poisons its CUDA context) with timeouts.
# usage: tools/gpu_tests.sh [files...]   (logs under gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
FILES=${@:-"tests/test_gpu_smplx.py tests/test_gpu_head.py tests/test_gpu_measure.py"}
for f in $FILES; do
  n=$(basename $f .py)
  echo "=== $f"
  timeout 600 python -m pytest $f -x -q -m gpu > gpurun_out/$n.log 2>&1
  echo "exit $?"; tail -n 12 gpurun_out/$n.log
done

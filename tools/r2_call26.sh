#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_zz_edges.py tests/test_gpu_smplx.py tests/test_gpu_zz_attributes.py tests/test_gpu_zz_metrics.py -q -m gpu 2>&1 | tail -25 | cut -c1-400

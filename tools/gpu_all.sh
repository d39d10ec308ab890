#!/bin/bash
# Full GPU suite as the driver runs it (single process), then smoke.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2

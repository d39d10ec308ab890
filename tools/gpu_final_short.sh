#!/bin/bash
# Round-end check without the profiler passes: full GPU suite, smoke, bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 2600 gpurun_out/bench_final.json

#!/bin/bash
# Round-end check: full GPU suite, smoke, bench (both arms), ncu launch list of one bench step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 3000 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 1200 gpurun_out/bench_ref.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches_final.csv python tools/profile_step.py 64 full 1 > gpurun_out/ncu_final.log 2>&1
tail -3 gpurun_out/ncu_final.log

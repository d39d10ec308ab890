#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_s8.json 2> gpurun_out/bench_s8.err
python -c "import json; l=json.loads(open('gpurun_out/bench_s8.json').read().strip().splitlines()[-1]); print('BENCH', l['value'], l['ms_per_step'], l['e2e'])"
tail -3 gpurun_out/bench_s8.err

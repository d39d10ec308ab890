#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2g4; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_n4c.json 2> $O/bench_n4c.err; echo "bench4 rc $?"
python - <<'P'
import json
l = [json.loads(x) for x in open('gpurun_out/r2g4/bench_n4c.json').read().strip().splitlines() if x.startswith('{')][-1]
print('N=4 value %.0f e2e %.0f (%.3f ms) step %.3f ms' % (l['value'], l['e2e']['value'], l['e2e']['ms_per_step'], l['ms_per_step']), 'config5', (l.get('config5') or {}).get('value'))
P

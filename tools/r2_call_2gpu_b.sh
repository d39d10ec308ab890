#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2g2b; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench2 rc $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref2 rc $?"
python - <<'P'
import json
l = [json.loads(x) for x in open('gpurun_out/r2g2b/bench_n2.json').read().strip().splitlines() if x.startswith('{')][-1]
print('N=2 value %.0f e2e %.0f step %.3f ms' % (l['value'], l['e2e']['value'], l['ms_per_step']), 'config5', (l.get('config5') or {}).get('value'))
r = [json.loads(x) for x in open('gpurun_out/r2g2b/bench_ref_n2.json').read().strip().splitlines() if x.startswith('{')]
print('ref lines', len(r), r[-1].get('value') if r else None)
P

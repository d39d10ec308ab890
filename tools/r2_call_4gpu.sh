#!/bin/bash
# 4-GPU run: the bench line at N = 4 (weak scaling, config5 scatter / gather) + the 2-GPU dist tests on the same box.
cd "$(dirname "$0")/.."
O=gpurun_out/r2g4; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader; nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench4 rc $?"; tail -3 $O/bench_n4.err | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_dist.py -q -m gpu > $O/dist.log 2>&1; echo "dist rc $?"; tail -2 $O/dist.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench1 rc $?"
python - <<'P'
import json
for n in (1, 4):
    try:
        l = [json.loads(x) for x in open(f'gpurun_out/r2g4/bench_n{n}.json').read().strip().splitlines() if x.startswith('{')][-1]
        print('N=%d value %.0f e2e %.0f ms %.3f' % (n, l['value'], l['e2e']['value'], l['ms_per_step']), 'config5', (l.get('config5') or {}).get('value'), l.get('clocks'))
    except Exception as e:
        print('ERR', n, e)
P

#!/bin/bash
# 2-GPU run: NCCL scatter / gather tests (fp32 and uint8 paths) and the bench line with config5.
cd "$(dirname "$0")/.."
O=gpurun_out/r2g2; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu > $O/dist.log 2>&1; echo "dist rc $?"; tail -4 $O/dist.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench2 rc $?"; tail -3 $O/bench_n2.err | cut -c1-300
python - <<'P'
import json
try:
    l = [json.loads(x) for x in open('gpurun_out/r2g2/bench_n2.json').read().strip().splitlines() if x.startswith('{')][-1]
    print('N=2 value %.0f e2e %.0f' % (l['value'], l['e2e']['value']), 'config5', l.get('config5'))
except Exception as e:
    print('ERR', e)
P

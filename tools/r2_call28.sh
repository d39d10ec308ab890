#!/bin/bash
# third A slice in the halo kernel: parity + A/B
cd "$(dirname "$0")/.."
O=gpurun_out/r2c28; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_conv.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300
run() { n=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$n.json 2> $O/bench_$n.err
  python - "$n" <<'P'
import json, sys
n = sys.argv[1]
try:
    l = json.loads(open(f'gpurun_out/r2c28/bench_{n}.json').read().strip().splitlines()[-1])
    print('%-12s value %.0f e2e %.0f step_ms %.3f hrnet_ms %.3f frac %.4f cfg2_ms %.3f clocks %s' % (n, l['value'], l['e2e']['value'], l['ms_per_step'], l['roofline']['ms'], l['roofline']['frac'], l['config2']['ms'], l['clocks']))
except Exception as e:
    print(n, 'ERR', e, open(f'gpurun_out/r2c28/bench_{n}.err').read()[-600:])
P
}
run warm X=1
run s3 X=1
run s2 SHAPY_CONV_ASLICES=2
run s3b X=1
run s2b SHAPY_CONV_ASLICES=2
SHAPY_CONV_DEBUG=1 timeout 120 python -c "
import torch
from shapy_b200 import synth
m = synth.build_synthetic_regressor().cuda().eval()
m.backbone(torch.zeros(64,3,224,224,device='cuda'))
" 2>&1 | grep "halo\]" | sort | uniq -c | sort -rn | head -30 > $O/halo_plans.txt; grep -c "slices in flight: 3" $O/halo_plans.txt

#!/bin/bash
# even grids / late PDL release A/B
cd "$(dirname "$0")/.."
O=gpurun_out/r2c13; mkdir -p $O
run() { n=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$n.json 2> $O/bench_$n.err
  python - "$n" <<'P'
import json, sys
n = sys.argv[1]
try:
    l = json.loads(open(f'gpurun_out/r2c13/bench_{n}.json').read().strip().splitlines()[-1])
    print('%-20s value %.0f e2e %.0f step_ms %.3f hrnet_ms %.3f frac %.4f cfg2_ms %.3f clocks %s' % (n, l['value'], l['e2e']['value'], l['ms_per_step'], l['roofline']['ms'], l['roofline']['frac'], l['config2']['ms'], l['clocks']))
except Exception as e:
    print(n, 'ERR', e, open(f'gpurun_out/r2c13/bench_{n}.err').read()[-600:])
P
}
run even1_late1 X=1
run even0_late0 SHAPY_CONV_EVENGRID=0 SHAPY_PDL_LATE=0
run even1_late0 SHAPY_PDL_LATE=0
run even0_late1 SHAPY_CONV_EVENGRID=0
run even1_late1_b X=1
run even0_late0_b SHAPY_CONV_EVENGRID=0 SHAPY_PDL_LATE=0
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_hrnet.py tests/test_gpu_e2e.py -q -m gpu 2>&1 | tail -3

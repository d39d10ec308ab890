#!/bin/bash
# per-role PDL wait A/B + conv parity
cd "$(dirname "$0")/.."
O=gpurun_out/r2c25; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_conv.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300
run() { n=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$n.json 2> $O/bench_$n.err
  python - "$n" <<'P'
import json, sys
n = sys.argv[1]
try:
    l = json.loads(open(f'gpurun_out/r2c25/bench_{n}.json').read().strip().splitlines()[-1])
    print('%-20s value %.0f e2e %.0f step_ms %.3f hrnet_ms %.3f frac %.4f cfg2_ms %.3f lbs %.1f us clocks %s' % (n, l['value'], l['e2e']['value'], l['ms_per_step'], l['roofline']['ms'], l['roofline']['frac'], l['config2']['ms'], l['roofline_lbs']['ms']*1e3, l['clocks']))
except Exception as e:
    print(n, 'ERR', e, open(f'gpurun_out/r2c25/bench_{n}.err').read()[-600:])
P
}
run late2 X=1
run late1 SHAPY_PDL_LATE=1
run late2b X=1
run late1b SHAPY_PDL_LATE=1

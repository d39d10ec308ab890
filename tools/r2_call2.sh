#!/bin/bash
# Round 2, GPU call 2: lane executor A/B (lanes 1/2/4, PDL policy 0/1/2, halo for 64-channel blocks), full GPU suite,
# batch sweep of the layer1 kernels (L2 residency question).
cd "$(dirname "$0")/.."
O=gpurun_out/r2c2; mkdir -p $O
timeout 1200 python -m pytest tests/ -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc $?"; tail -8 $O/pytest.log | cut -c1-300
run() { # name, env...
  n=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$n.json 2> $O/bench_$n.err
  python - "$n" <<'P'
import json, sys
n = sys.argv[1]
try:
    l = json.loads(open(f'gpurun_out/r2c2/bench_{n}.json').read().strip().splitlines()[-1])
    print('%-16s value %.0f e2e %.0f step_ms %.3f hrnet_ms %.3f frac %.4f clocks %s' % (n, l['value'], l['e2e']['value'], l['ms_per_step'], l['roofline']['ms'], l['roofline']['frac'], l['clocks']))
except Exception as e:
    print(n, 'ERR', e, open(f'gpurun_out/r2c2/bench_{n}.err').read()[-600:])
P
}
run lanes4 SHAPY_HRNET_LANES=4
run lanes1 SHAPY_HRNET_LANES=1
run lanes1_pdl1 SHAPY_HRNET_LANES=1 SHAPY_PDL=1
run lanes4_pdl1 SHAPY_HRNET_LANES=4 SHAPY_PDL=1
run lanes4_pdl0 SHAPY_HRNET_LANES=4 SHAPY_PDL=0
run lanes2 SHAPY_HRNET_LANES=2
run lanes4_halo64 SHAPY_HRNET_LANES=4 SHAPY_CONV_HALO_MAXKCH=64
run lanes4_b32_fp16 SHAPY_HRNET_LANES=4 BENCH_ARGS=1
for B in 8 16 32; do
  SHAPY_CONV_HALO_MAXKCH=64 timeout 300 python tools/conv_layer_bench.py $B 1 b64,b1x1a,b1x1b,t48,c48 2>&1 | grep conv_test
done

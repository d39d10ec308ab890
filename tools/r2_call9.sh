#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c9; mkdir -p $O
run() { n=$1; shift; env "$@" timeout 200 python tools/e2e_repro.py 64 8 u8 > $O/repro_$n.log 2>&1; echo "$n rc $? ok=$(grep -c '^rep' $O/repro_$n.log) $(grep -m1 -i 'error' $O/repro_$n.log | cut -c1-120)"; }
run a X=1; run b X=1; run c X=1
timeout 1200 python -m pytest tests/ -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc $?"; tail -6 $O/pytest.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc $?"; tail -3 $O/bench.err
python - <<'P'
import json
try:
    l = json.loads(open('gpurun_out/r2c9/bench.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f step_ms %.3f hrnet_ms %.3f frac %.4f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['roofline']['ms'], l['roofline']['frac']))
    for k in ('roofline_lbs', 'roofline_shape', 'roofline_measure', 'config2', 'cpu_baseline', 'selfcheck'):
        print(k, l.get(k))
except Exception as e:
    print('ERR', e)
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"smplx_lbs|smplx_joints" -c 4 -o $O/lbs_full -f \
  python tools/profile_step.py 64 lbs 1 > $O/ncu_lbs.log 2>&1; echo "ncu rc $?"
timeout 600 python tools/gpu_baselines.py > $O/gpu_baselines.json 2> $O/gpu_baselines.err; echo "baselines rc $?"; cat $O/gpu_baselines.json | head -40; tail -3 $O/gpu_baselines.err

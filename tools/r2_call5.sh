#!/bin/bash
# Round 2, GPU call 5: full GPU suite (fused LBS, measure v2, A2B, v2v, uint8 pipeline), the new bench line, ncu of the LBS /
# measurement kernels.
cd "$(dirname "$0")/.."
O=gpurun_out/r2c5; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_measure.py -q -m gpu > $O/measure.log 2>&1; echo "measure rc $?"; tail -12 $O/measure.log | cut -c1-400
timeout 1200 python -m pytest tests/ -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc $?"; tail -12 $O/pytest.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc $?"; tail -5 $O/bench.err
python - <<'P'
import json
try:
    l = json.loads(open('gpurun_out/r2c5/bench.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f step_ms %.3f hrnet_ms %.3f frac %.4f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['roofline']['ms'], l['roofline']['frac']))
    for k in ('roofline_lbs', 'roofline_shape', 'roofline_measure', 'config2', 'cpu_baseline', 'selfcheck'):
        print(k, l.get(k))
except Exception as e:
    print('ERR', e)
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"smplx_|measure_" -c 8 -o $O/lbs_full -f \
  python tools/profile_step.py 64 lbs 1 > $O/ncu_lbs.log 2>&1; echo "ncu rc $?"; tail -2 $O/ncu_lbs.log

#!/bin/bash
# Round 2, GPU call 4: first run of the fused tcgen05 LBS kernel and the measure v2 kernel (each group of tests in its
# own process under a timeout: a hang must not take the box), then the suite and the bench.
cd "$(dirname "$0")/.."
O=gpurun_out/r2c4; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_smplx.py -q -m gpu -x > $O/smplx.log 2>&1; echo "smplx rc $?"; tail -15 $O/smplx.log | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_measure.py -q -m gpu > $O/measure.log 2>&1; echo "measure rc $?"; tail -12 $O/measure.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "benchmark_batch" > $O/conv.log 2>&1; echo "conv rc $?"; tail -8 $O/conv.log | cut -c1-300
timeout 900 python -m pytest tests/ -q -m gpu --deselect tests/test_gpu_conv.py > $O/pytest.log 2>&1; echo "pytest rc $?"; tail -8 $O/pytest.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc $?"
python - <<'P'
import json
try:
    l = json.loads(open('gpurun_out/r2c4/bench.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f step_ms %.3f hrnet_ms %.3f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['roofline']['ms']))
    for k in ('roofline_lbs', 'roofline_shape', 'measure_4096'):
        print(k, l[k])
except Exception as e:
    print('ERR', e, open('gpurun_out/r2c4/bench.err').read()[-800:])
P
SHAPY_LBS_FUSED=0 SHAPY_MEASURE_V1=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_old.json 2> $O/bench_old.err
python - <<'P'
import json
try:
    l = json.loads(open('gpurun_out/r2c4/bench_old.json').read().strip().splitlines()[-1])
    for k in ('roofline_lbs', 'measure_4096'):
        print('old', k, l[k])
except Exception as e:
    print('ERR', e)
P

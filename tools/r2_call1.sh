#!/bin/bash
# Round 2, GPU call 1: state check (full GPU suite incl. the new benchmarked-configuration parity tests), baseline bench,
# halo-for-64-channel A/B, per-layer timings + phase counters, ncu --set full of the current conv kernels.
cd "$(dirname "$0")/.."
O=gpurun_out/r2c1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests/ -q -m gpu -x > $O/pytest.log 2>&1; echo "pytest rc $?"; tail -5 $O/pytest.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_base.json 2> $O/bench_base.err; echo "bench rc $?"
SHAPY_CONV_HALO_MAXKCH=64 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_halo64.json 2> $O/bench_halo64.err; echo "bench halo64 rc $?"
python - <<'P'
import json
for n in ('base','halo64'):
    try:
        l=json.loads(open(f'gpurun_out/r2c1/bench_{n}.json').read().strip().splitlines()[-1])
        print(n, 'value %.0f e2e %.0f hrnet_ms %.3f lbs_ms %.4f' % (l['value'], l['e2e']['value'], l['roofline']['ms'], l['roofline_lbs']['ms']))
    except Exception as e: print(n, 'ERR', e)
P
L=c48,c96,c192,c384,b64,b1x1b,d96,d48,s2_192_384
SHAPY_CONV_DEBUG=1 timeout 300 python tools/conv_layer_bench.py 64 1 $L > $O/layers_base.txt 2>&1
SHAPY_CONV_HALO_MAXKCH=64 SHAPY_CONV_DEBUG=1 timeout 300 python tools/conv_layer_bench.py 64 1 c192,c384,b64,s2_192_384 > $O/layers_halo64.txt 2>&1
grep conv_test $O/layers_base.txt; echo ---; grep conv_test $O/layers_halo64.txt
SHAPY_CONV_PHASES=1 SHAPY_CONV_TEST_REPS=3 timeout 300 python tools/conv_layer_bench.py 64 1 c48,c96 > $O/phases_base.txt 2>&1
SHAPY_CONV_HALO_MAXKCH=64 SHAPY_CONV_PHASES=1 SHAPY_CONV_TEST_REPS=3 timeout 300 python tools/conv_layer_bench.py 64 1 c192,c384 > $O/phases_halo64.txt 2>&1
grep phases $O/phases_base.txt | tail -8; grep phases $O/phases_halo64.txt | tail -8
SHAPY_CONV_TEST_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -c 14 -o $O/conv_full -f \
  python tools/conv_layer_bench.py 64 1 c48,c96,c192,c384,b64,b1x1b,d96 > $O/ncu_conv.log 2>&1; echo "ncu rc $?"; tail -2 $O/ncu_conv.log
ls -la $O

#!/bin/bash
# Round 2 evidence run (1 GPU): final bench (both arms), launch list + DRAM traffic of one step, ncu --set full of the
# fp16-mode conv kernels (configs[1]) and of the measurement / shape kernels.
cd "$(dirname "$0")/.."
O=gpurun_out/r2ev; mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc $?"; tail -2 $O/bench_final.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc $?"
SHAPY_HRNET_GRAPH=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file $O/launches_step.csv python tools/profile_step.py 64 full 1 > $O/ncu_step.log 2>&1; echo "launch list rc $?"
SHAPY_HRNET_GRAPH=0 timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none -c 1500 --csv --log-file $O/traffic.csv python tools/profile_step.py 64 hrnet 1 > $O/ncu_traffic.log 2>&1; echo "traffic rc $?"
SHAPY_CONV_TEST_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -c 10 -o $O/conv_fp16_full -f \
  python tools/conv_layer_bench.py 32 0 c48,c96,c192,c384,b1x1b > $O/ncu_fp16.log 2>&1; echo "ncu fp16 rc $?"
SHAPY_HRNET_LANES=1 SHAPY_HRNET_GRAPH=0 timeout 300 python tools/hrnet_trace.py 64 $O/trace_lanes1.txt > $O/trace_summary1.txt 2>&1
python - <<'P'
import json
for n in ('final', 'ref'):
    try:
        l = json.loads(open(f'gpurun_out/r2ev/bench_{n}.json').read().strip().splitlines()[-1])
        print(n, 'value %.0f e2e %.0f' % (l['value'], l['e2e']['value']), {k: round(v['frac'], 4) for k, v in l.items() if isinstance(v, dict) and 'frac' in v})
    except Exception as e:
        print(n, 'ERR', e)
P

#!/bin/bash
# Round 2 final evidence (1 GPU): full GPU suite, bench (both arms), launch list of one step, ncu --set full of the fused
# LBS kernel, LBS timeline.
cd "$(dirname "$0")/.."
O=gpurun_out/r2final; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"; tail -2 $O/bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc $?"
SHAPY_HRNET_GRAPH=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file $O/launches_step.csv python tools/profile_step.py 64 full 1 > $O/ncu_step.log 2>&1; echo "launch list rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:smplx_ --launch-skip 6 --launch-count 2 -f -o $O/lbs_full \
  python tools/lbs_time.py 64 > $O/ncu_lbs.log 2>&1; echo "ncu lbs rc $?"
{ echo "# tools/lbs_time.py (event-timed, L2 flushed by a 256 MB write before every call; LBS + joints kernels)";
  echo "## fused kernel, skinning-transform blend on tcgen05 (default)"; timeout 120 python tools/lbs_time.py 64 256 4096 2>&1 | tail -3;
  echo "## fused kernel, skinning from shared-memory A_j (SHAPY_LBS_TV=0)"; SHAPY_LBS_TV=0 timeout 120 python tools/lbs_time.py 64 256 4096 2>&1 | tail -3;
  echo "## three-kernel path (SHAPY_LBS_FUSED=0)"; SHAPY_LBS_FUSED=0 timeout 120 python tools/lbs_time.py 64 256 4096 2>&1 | tail -3;
  echo "## per-role cycle stamps, B = 64"; SHAPY_LBS_DEBUG=1 timeout 60 python tools/lbs_time.py 64 2>&1 | grep "lbs\]" | tail -20;
  echo "## per-role cycle stamps, B = 4096 (items 8 and 9 of every CTA)"; SHAPY_LBS_DEBUG=1 timeout 60 python tools/lbs_time.py 4096 2>&1 | grep "lbs\]" | tail -20; } > $O/lbs_timeline.txt 2>&1
python - <<'P'
import json
for n in ('n1', 'ref'):
    try:
        l = json.loads(open(f'gpurun_out/r2final/bench_{n}.json').read().strip().splitlines()[-1])
        print(n, 'value %.0f e2e %.0f' % (l['value'], l['e2e']['value']), {k: round(v['frac'], 4) for k, v in l.items() if isinstance(v, dict) and 'frac' in v}, l.get('clocks'))
    except Exception as e:
        print(n, 'ERR', e)
P

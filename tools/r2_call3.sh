#!/bin/bash
# Round 2, GPU call 3: CUDA-graph replay A/B on the lane executor, lane timeline trace, GPU suite.
cd "$(dirname "$0")/.."
O=gpurun_out/r2c3; mkdir -p $O
run() { n=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$n.json 2> $O/bench_$n.err
  python - "$n" <<'P'
import json, sys
n = sys.argv[1]
try:
    l = json.loads(open(f'gpurun_out/r2c3/bench_{n}.json').read().strip().splitlines()[-1])
    print('%-16s value %.0f e2e %.0f step_ms %.3f hrnet_ms %.3f frac %.4f launches %d clocks %s' % (n, l['value'], l['e2e']['value'], l['ms_per_step'], l['roofline']['ms'], l['roofline']['frac'], l['gpu_launches'], l['clocks']))
except Exception as e:
    print(n, 'ERR', e, open(f'gpurun_out/r2c3/bench_{n}.err').read()[-800:])
P
}
run graph SHAPY_CONV_DEBUG=1
grep "hrnet\]" $O/bench_graph.err | head -5
run nograph SHAPY_HRNET_GRAPH=0
run graph_halo64 SHAPY_CONV_HALO_MAXKCH=64
run graph2 SHAPY_X=1
run nograph_pdl2 SHAPY_HRNET_GRAPH=0 SHAPY_PDL=2
timeout 300 python tools/hrnet_trace.py 64 $O/trace_lanes4.txt > $O/trace_summary.txt 2>&1; tail -60 $O/trace_summary.txt
SHAPY_HRNET_LANES=1 timeout 300 python tools/hrnet_trace.py 64 $O/trace_lanes1.txt > $O/trace_summary1.txt 2>&1; head -3 $O/trace_summary1.txt
timeout 1200 python -m pytest tests/ -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc $?"; tail -6 $O/pytest.log | cut -c1-300

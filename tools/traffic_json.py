"""DRAM bytes of the conv launches of one HRNet forward from an ncu --metrics dram__bytes_*.sum,gpu__time_duration.sum csv.
usage: python tools/traffic_json.py gpurun_out/traffic.csv "description" > profiles/rNN_traffic.json"""
import io, json, sys
import pandas as pd
lines = open(sys.argv[1]).read().splitlines()
i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
df = pd.read_csv(io.StringIO('\n'.join(lines[i:])))
df['v'] = df['Metric Value'].astype(str).str.replace(',', '').astype(float)
conv = df[df['Kernel Name'].astype(str).str.contains('conv_')]
unit = {r['Metric Name']: r['Metric Unit'] for _, r in conv.iterrows()}


def total(metric):
    d = conv[conv['Metric Name'] == metric]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1.0, 'us': 1e3, 'usecond': 1e3, 'msecond': 1e6}
    return float(sum(v * scale.get(u, 1.0) for v, u in zip(d['v'], d['Metric Unit'])))
rd, wr, ns = total('dram__bytes_read.sum'), total('dram__bytes_write.sum'), total('gpu__time_duration.sum')
print(json.dumps({'what': 'DRAM bytes (read+write) of all conv_* launches of one HRNet forward, B=64, 224x224, split mode; ncu --metrics '
                          'dram__bytes_read.sum,dram__bytes_write.sum (cold-cache, serialised replays: reads that hit the 126 MB L2 in a '
                          'real step come from DRAM here)',
                  'conv_launches': int(len(conv[conv['Metric Name'] == 'gpu__time_duration.sum'])), 'conv_dram_bytes_read': rd,
                  'conv_dram_bytes_write': wr, 'conv_dram_bytes_per_step': rd + wr, 'conv_us_per_step_ncu': ns / 1e3,
                  'source': sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}, indent=1))

"""Joins an ncu launch list of one HRNet forward (tools/profile_step.py ... hrnet) with the op program, so every
launch gets its layer shape and FLOPs.  usage: python tools/join_launches.py gpurun_out/launches.csv [batch]"""
import io, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandas as pd
from shapy_b200 import synth

path = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lines = open(path).read().splitlines()
i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
df = pd.read_csv(io.StringIO('\n'.join(lines[i:])))
df = df[df['Metric Name'] == 'gpu__time_duration.sum']
df['us'] = df['Metric Value'].astype(str).str.replace(',', '').astype(float) / 1000
names = df['Kernel Name'].astype(str).tolist()
us = df['us'].tolist()
keep = [(n, t) for n, t in zip(names, us) if any(s in n for s in ('conv_', 'stem', 'fuse_kernel', 'pool_kernel'))]
model = synth.build_synthetic_regressor()
convs, ops, slots, feat, layer_slots = model.backbone.build_program()
assert len(keep) == len(ops), (len(keep), len(ops))
rows = []
for (n, t), o in zip(keep, ops):
    kind = o['kind']
    if kind in (0, 1):
        c = convs[o['conv']]
        div = slots[o['out_slot']]['div']
        H = 224 // div
        fl = 2.0 * B * H * H * c['cout'] * c['cin'] * c['ksize'] ** 2
        key = f"{c['cin']}->{c['cout']} k{c['ksize']} s{c['stride']} out{H}x{H}" + (' +res' if o['res_slot'] >= 0 else '')
        rows.append((key, n.split('(')[0].replace('void shapy::', '')[:28], t, fl))
    else:
        rows.append(('fuse' if kind == 2 else 'pool', n.split('(')[0][:28], t, 0.0))
r = pd.DataFrame(rows, columns=['layer', 'kernel', 'us', 'flops'])
g = r.groupby(['layer', 'kernel']).agg(n=('us', 'count'), us_total=('us', 'sum'), us_mean=('us', 'mean'), flops=('flops', 'mean'))
g['TF/s'] = g['flops'] / g['us_mean'] * 1e-6
g = g.sort_values('us_total', ascending=False)
pd.set_option('display.width', 200)
print(g.to_string(float_format=lambda v: f'{v:.1f}'))
print('total us', r.us.sum())

"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line: instructions executed and
stall samples.  usage: python tools/ncu_lines.py src.csv [by=inst|samp] [n]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
by = sys.argv[2] if len(sys.argv) > 2 else 'inst'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
hdr = rows[2]
iSamp = hdr.index('# Samples'); iInst = hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
def f(x):
    try: return int(x)
    except ValueError: return 0
agg = collections.OrderedDict(); ti = ts = 0
for r in rows[3:]:
    if r[0] == '' or not r[0].isdigit(): continue
    k = (int(r[0]), r[1].strip()[:90])
    a = agg.setdefault(k, [0, 0, collections.Counter()])
    a[0] += f(r[iInst]); a[1] += f(r[iSamp]); ti += f(r[iInst]); ts += f(r[iSamp])
    for i in stall_cols:
        v = f(r[i])
        if v: a[2][hdr[i][6:]] += v
print(f'total warp instructions {ti / 1e6:.1f}M, samples {ts}')
key = (lambda kv: -kv[1][0]) if by == 'inst' else (lambda kv: -kv[1][1])
for (ln, src), (i, s, c) in sorted(agg.items(), key=key)[:n]:
    print(f'{ln:5d} inst {i / 1e6:7.2f}M {100 * i / ti:5.1f}%  samp {100 * s / ts:5.1f}% {dict(c.most_common(3))} | {src}')

#!/usr/bin/env python
"""Summarises .ncu-rep captures into a markdown table (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/rNN_ncu_summary.md"""
import csv
import io
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_%'),
        ('launch__registers_per_thread', 'regs'), ('launch__shared_mem_per_block_dynamic', 'dsmem'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'block')]


def main():
    print('| report | kernel | ' + ' | '.join(n for _, n in WANT) + ' |')
    print('|---|---|' + '---|' * len(WANT))
    for rep in sys.argv[1:]:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[idx['Kernel Name']].replace('void ', '').split('(')[0][:48]
            cells = []
            for key, _ in WANT:
                if key in idx:
                    v = r[idx[key]]
                    u = units[idx[key]]
                    try:
                        v = f'{float(v.replace(",", "")):.4g}'
                    except ValueError:
                        pass
                    cells.append(f'{v} {u}'.strip())
                else:
                    cells.append('-')
            print(f'| {rep.split("/")[-1]} | {name} | ' + ' | '.join(cells) + ' |')


if __name__ == '__main__':
    main()

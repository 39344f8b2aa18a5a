#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total
time, share of the captured window. Usage: summarize_launches.py launches.csv [last N launches] > summary.md"""
import csv
import re
import sys
from collections import defaultdict


def main(path, last=None):
    rows = []
    with open(path, newline='') as fh:
        lines = [ln for ln in fh if not ln.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        val = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1e-6)
        name = r['Kernel Name'].replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
        name = re.sub(r'^void ', '', name)
        m = re.match(r'([\w:]+)(<[^(]*>)?', name)
        name = (m.group(1) + (m.group(2) or '')) if m else name
        name = name if len(name) < 70 else name[:67] + '...'
        rows.append((name, val * scale))
    if last:
        rows = rows[-int(last):]        # e.g. the launches of the final (eager, profiled) step of bench.py
    tot = sum(t for _, t in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, t in rows:
        agg[n][0] += 1
        agg[n][1] += t
    print('| kernel | launches | total ms | share | avg us |')
    print('|---|---:|---:|---:|---:|')
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.2f | %.1f %% | %.1f |' % (n, c, t, 100 * t / tot, 1e3 * t / c))
    print('| **total** | %d | %.2f | 100 %% | |' % (len(rows), tot))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)

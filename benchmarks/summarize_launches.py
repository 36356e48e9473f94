#!/usr/bin/env python
"""Summarise an ncu launch list (``--metrics gpu__time_duration.sum --csv --log-file X``)
into per-kernel totals:  python benchmarks/summarize_launches.py X.csv [out.json] [skip_first_N]"""
import csv
import json
import re
import sys


def main():
    path = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else None
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if 'Kernel Name' in r:
                hdr = r
            continue
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        try:
            v = float(d['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        unit = d.get('Metric Unit', 'ns')
        us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
        rows.append((d['Kernel Name'], us))
    rows = rows[skip:]
    agg = {}
    for name, us in rows:
        key = re.sub(r'\(.*$', '', name)[:100]
        a = agg.setdefault(key, {'n': 0, 'us': 0.0})
        a['n'] += 1
        a['us'] += us
    total = sum(a['us'] for a in agg.values())
    items = sorted(agg.items(), key=lambda kv: -kv[1]['us'])
    res = {'total_us': round(total, 3), 'launches': len(rows),
           'kernels': {k: {'n': v['n'], 'us': round(v['us'], 3), 'pct': round(100 * v['us'] / max(total, 1e-9), 2)}
                       for k, v in items}}
    txt = json.dumps(res, indent=1)
    if out:
        with open(out, 'w') as f:
            f.write(txt)
    for k, v in items[:40]:
        print('%9.1f us %5.1f%% n=%4d  %s' % (v['us'], 100 * v['us'] / total, v['n'], k))
    print('total %.1f us, %d launches' % (total, len(rows)))


if __name__ == '__main__':
    main()

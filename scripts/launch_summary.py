#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of ONE bench step
(the launches between two consecutive k_lm_finish).  usage: launch_summary.py launches.csv [step_index_from_end]"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = []
for x in csv.DictReader(lines):
    if x.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(x['Metric Value'].replace(',', '')); u = x['Metric Unit']
    v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
    rows.append((x['Kernel Name'].split('(')[0].replace('lisreg::', ''), v))
fin = [i for i, r in enumerate(rows) if r[0] == 'k_lm_finish']
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
a, b = fin[-k - 1] + 1, fin[-k] + 1
agg = collections.OrderedDict()
for n, v in rows[a:b]:
    d = agg.setdefault(n[:60], [0, 0.0]); d[0] += 1; d[1] += v
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.3f | %.1f |" % (n, v[0], v[1], v[1] / tot, v[1] / v[0]))
print("| total | %d | %.1f | 1 | |" % (sum(v[0] for v in agg.values()), tot))

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
# steps = launch ranges between consecutive k_lm_finish; default: the longest one (a full device-resident batch,
# not one chunk of the pipelined e2e path)
steps = [(fin[i] + 1, fin[i + 1] + 1) for i in range(len(fin) - 1)]
if len(sys.argv) > 2 and sys.argv[2] == "median":      # a typical step (streaming: a frame without a map rebuild)
    a, b = sorted(steps, key=lambda ab: sum(v for _, v in rows[ab[0]:ab[1]]))[len(steps) // 2]
elif len(sys.argv) > 2:
    a, b = steps[-int(sys.argv[2])]
else:
    a, b = max(steps, key=lambda ab: sum(v for _, v in rows[ab[0]:ab[1]]))
agg = collections.OrderedDict()
for n, v in rows[a:b]:
    d = agg.setdefault(n[:60], [0, 0.0]); d[0] += 1; d[1] += v
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.3f | %.1f |" % (n, v[0], v[1], v[1] / tot, v[1] / v[0]))
print("| total | %d | %.1f | 1 | |" % (sum(v[0] for v in agg.values()), tot))
for kn in ("k_lm_knn", "k_lm_resid"):
    print("%s per iteration (us): %s" % (kn, [round(v) for n, v in rows[a:b] if n == kn]))

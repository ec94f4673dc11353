#!/usr/bin/env python
"""profiles/lm_iter_traffic.json from an ncu metrics CSV (dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum over the kNN / LM kernels of one bench step): DRAM bytes per Gauss-Newton iteration.
usage: make_traffic.py metrics.csv out.json"""
import collections, csv, json, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
per = collections.OrderedDict()
for x in csv.DictReader(lines):
    k = x['ID']; d = per.setdefault(k, {"name": x['Kernel Name'].split('(')[0].replace('lisreg::', '').replace('void ', '')})
    v = float(x['Metric Value'].replace(',', '')); u = x['Metric Unit']
    if x['Metric Name'].startswith('dram__bytes'):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    elif x['Metric Name'] == 'gpu__time_duration.sum':
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3}[u]
    d[x['Metric Name']] = v
rows = list(per.values())
# one step = the launches from the first k_knn_check after a k_lm_init to the last k_lm_solve before k_lm_finish
names = [r["name"] for r in rows]
start = names.index("k_lm_init") + 1
end = names.index("k_lm_finish", start)
step = rows[start:end]
iters = sum(1 for r in step if r["name"] == "k_lm_solve")
agg = collections.OrderedDict()
for r in step:
    a = agg.setdefault(r["name"], {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
    a["launches"] += 1; a["dram_bytes"] += r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0); a["us"] += r.get("gpu__time_duration.sum", 0)
tot = sum(a["dram_bytes"] for a in agg.values())
out = {"dram_bytes_per_iteration": tot / iters, "iterations": iters, "per_kernel_per_iteration": {k: {"dram_bytes": v["dram_bytes"] / iters, "us": v["us"] / iters, "launches": v["launches"] / iters} for k, v in agg.items()},
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, python bench.py --steps 1 --warmup 1 --no-cpu (256 frames per step)"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))

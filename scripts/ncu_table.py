#!/usr/bin/env python
"""Markdown table of the key metrics of every launch in an `ncu --page raw --csv` export.  usage: ncu_table.py raw.csv"""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
hdr, units = r[0], r[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [('gpu__time_duration.sum', 'dur us'), ('dram__bytes_read.sum', 'dram rd MB'), ('dram__bytes_write.sum', 'dram wr MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'), ('launch__registers_per_thread', 'regs'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps act %'), ('smsp__thread_inst_executed_per_inst_executed.ratio', 'lanes/inst'),
        ('smsp__inst_executed.sum', 'warp inst M'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
        ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'), ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_sb'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_sb'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier'), ('launch__grid_size', 'grid')]
print('| kernel | ' + ' | '.join(c[1] for c in cols) + ' |')
print('|' + '---|' * (len(cols) + 1))
for row in r[2:]:
    name = row[ix['Kernel Name']].split('(')[0].replace('lisreg::', '').replace('void ', '')
    vals = []
    for c, _ in cols:
        v, u = row[ix[c]], units[ix[c]]
        try:
            f = float(v.replace(',', ''))
            if c.startswith('dram__bytes'):
                f *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}[u]
            if c == 'gpu__time_duration.sum':
                f *= {'ns': 1e-3, 'us': 1, 'ms': 1e3}[u]
            if c == 'smsp__inst_executed.sum':
                f /= 1e6
            vals.append('%.1f' % f if abs(f) < 1e5 else '%.0f' % f)
        except ValueError:
            vals.append(v)
    print('| ' + name + ' | ' + ' | '.join(vals) + ' |')

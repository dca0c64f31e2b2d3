#!/usr/bin/env python
"""Per-launch list of ONE step from an ncu `--metrics gpu__time_duration.sum --csv` log: the launches between the 2nd and 3rd
k_prefilter (cold-cache, serialised times: compare shares).  Usage: python profiles/step_launches.py launches.csv [--agg]"""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, x in enumerate(rows) if 'k_prefilter' in x['Kernel Name']]
a, b = idx[1], idx[2]
tot, agg = 0.0, {}
for x in rows[a:b]:
    name = re.sub(r'\(.*', '', x['Kernel Name']).replace('void ', '')[:60]
    t = float(x['Metric Value'].replace(',', '')) / (1000 if x['Metric Unit'] == 'ns' else 1)
    tot += t
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1; agg[name][1] += t
    if '--agg' not in sys.argv:
        print(f"{t:9.1f} {x['Grid Size']:>14} {name}")
print(f"total us {tot:.1f} launches {b - a}")
if '--agg' in sys.argv:
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.1f} us {v[0]:4d} {100 * v[1] / tot:6.1f}%  {k}")

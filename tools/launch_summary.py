#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6}.get(r['Metric Unit'], 1.0)
    k = r['Kernel Name'].split('(')[0]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total device time {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:64]:64s} n={v[0]:6d} total={v[1] / 1e6:9.3f} ms avg={v[1] / v[0] / 1e3:9.1f} us share={v[1] / tot:.3f}")

"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python profiles/summarize_launches.py file.csv"""
import collections
import csv
import io
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(io.StringIO(''.join(lines)))
hdr = next(r)
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    name = row[ki].split('(')[0].replace('<unnamed>::', '').replace('void ', '')[:48]
    v = float(row[vi].replace(',', ''))
    v = v / 1e6 if row[ui] == 'ns' else (v / 1e3 if row[ui] == 'us' else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':50s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>10s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} {v[0]:8d} {v[1]:10.3f} {v[1] / tot * 100:6.1f}% {v[1] / v[0] * 1e3:10.1f}")
print(f"{'total':50s} {sum(v[0] for v in agg.values()):8d} {tot:10.3f}")

"""Aggregate an ncu launch list per kernel:  python profiles/summarize_launches.py file.csv
The list comes from `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --clock-control none --csv`
(one row per launch and metric); DRAM columns are printed when present."""
import collections
import csv
import io
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(io.StringIO(''.join(lines)))
hdr = next(r)
ki, mi, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])        # launches, ms, dram read B, dram write B
has_dram = False
for row in r:
    name = row[ki].split('(')[0].replace('<unnamed>::', '').replace('void ', '').replace('m3::', '')[:56]
    v = float(row[vi].replace(',', ''))
    u = row[ui]
    if row[mi].startswith('gpu__time_duration'):
        agg[name][0] += 1
        agg[name][1] += v / 1e6 if u in ('ns', 'nsecond') else (v / 1e3 if u in ('us', 'usecond') else (v * 1e3 if u in ('s', 'second') else v))
    else:
        has_dram = True
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
        agg[name][2 if 'read' in row[mi] else 3] += v * scale
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>10s}" + (f" {'dram rd GB':>11s} {'dram wr GB':>11s}" if has_dram else ''))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:58s} {v[0]:8d} {v[1]:10.3f} {v[1] / tot * 100:6.1f}% {v[1] / v[0] * 1e3:10.1f}"
          + (f" {v[2] / 1e9:11.3f} {v[3] / 1e9:11.3f}" if has_dram else ''))
print(f"{'total':58s} {sum(v[0] for v in agg.values()):8d} {tot:10.3f}")

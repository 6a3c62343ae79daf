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

# optional:  python profiles/summarize_launches.py file.csv --traffic-json profiles/ncu_traffic.json --candidates 16 [--source name]
# writes the DRAM traffic record bench.py's roofline.traffic is computed from (stamped with the kernel sources' sha)
if "--traffic-json" in sys.argv and has_dram:
    import hashlib, json, os, subprocess
    out = sys.argv[sys.argv.index("--traffic-json") + 1]
    cand = int(sys.argv[sys.argv.index("--candidates") + 1])
    src = sys.argv[sys.argv.index("--source") + 1] if "--source" in sys.argv else sys.argv[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()
    for f in ("zgemm.cu", "zgemm.cuh", "lu.cu", "lu.cuh"):
        h.update(open(os.path.join(root, "adaptive-matrix-solver_b200", "csrc", f), "rb").read())
    gemm = [v for k, v in agg.items() if k.startswith("zgemm3m")]
    rd = sum(v[2] for v in gemm); wr = sum(v[3] for v in gemm); n_l = sum(v[0] for v in gemm)
    try:
        commit = subprocess.run(["git", "-C", root, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        commit = None
    json.dump({"source": src, "commit": commit, "candidates_measured": cand, "gemm_launches": n_l,
               "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_candidate": (rd + wr) / cand,
               "kernel_sources_sha16": h.hexdigest()[:16],
               "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the zgemm3m_dmma_kernel launches of ONE generation"},
              open(out, "w"), indent=1)
    print(f"wrote {out}: {(rd + wr) / cand / 1e9:.3f} GB per candidate over {n_l} GEMM launches")

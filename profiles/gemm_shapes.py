"""Aggregate the per-launch GEMM log of the LU (MAUS_GEMM_LOG=1 python profiles/prof_step.py C steps 2> log) by shape class.
   python profiles/gemm_shapes.py log [peak_tflops]   -> time share and algorithmic TFLOP/s per class"""
import collections
import re
import sys

peak = float(sys.argv[2]) if len(sys.argv) > 2 else 37.1
rows = collections.defaultdict(lambda: [0, 0.0, 0.0])
for line in open(sys.argv[1]):
    m = re.match(r"gemm M (\d+) N (\d+) K (\d+) batch (\d+) ms ([\d.]+)", line)
    if not m:
        continue
    M, N, K, b, ms = int(m[1]), int(m[2]), int(m[3]), int(m[4]), float(m[5])
    if K >= 256:
        cls = f"bulk K={K}"
    elif M <= 128 and K == M:
        cls = f"U12 = L11^-1 A12 (M=K={K})"
    elif M <= 128:
        cls = f"row block (b) M={M}"
    else:
        cls = f"update K={K} N<={64 if N <= 64 else 384}"
    r = rows[cls]
    r[0] += 1; r[1] += ms; r[2] += 8.0 * M * N * K * b
tot = sum(r[1] for r in rows.values())
print(f"{'class':38s} {'launches':>8s} {'ms':>9s} {'share':>6s} {'TFLOP/s (alg)':>13s} {'of peak':>8s}")
for cls, (cnt, ms, fl) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    tf = fl / ms / 1e9
    print(f"{cls:38s} {cnt:8d} {ms:9.3f} {ms / tot:6.1%} {tf:13.2f} {tf / peak:8.1%}")
fl = sum(r[2] for r in rows.values())
print(f"{'total':38s} {sum(r[0] for r in rows.values()):8d} {tot:9.3f} {1:6.1%} {fl / tot / 1e9:13.2f} {fl / tot / 1e9 / peak:8.1%}")

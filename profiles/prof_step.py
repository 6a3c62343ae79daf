"""Small fixed workload for ncu: a few fused generations (resident vectors) at n=4096.  Usage:
   python profiles/prof_step.py [C] [steps]      (never a bench number: run under the profiler only)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaptive_matrix_solver_b200 as pkg                      # noqa: E402
from adaptive_matrix_solver_b200 import _abi                   # noqa: E402
from adaptive_matrix_solver_b200.workloads import k2_matrix, initial_vectors   # noqa: E402

C_ = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
A = k2_matrix(n)
V = initial_vectors(C_, n)
eng = pkg.MausEngine(0)
eng.set_matrix(A)
eng.upload_vectors(V)
alpha = np.full(C_, 0.01); psi = np.full(C_, 1e-20)
eng.step(_abi.EIGENVALUE, alpha, psi, V=None, rng_key=np.arange(C_, dtype=np.uint64))   # warm-up (allocations)
if not os.environ.get('NOPROF'):           # NOPROF=1: wall-clock per step only (the two-stream LU is disabled while profiling)
    eng.profile_reset(True)
for s in range(steps):
    t0 = time.perf_counter()
    out = eng.step(_abi.EIGENVALUE, alpha, psi, V=None, rng_key=np.arange(C_, dtype=np.uint64) + np.uint64(1000 * s))
    print(f"step {s}: {1e3 * (time.perf_counter() - t0):.1f} ms, launches so far {eng.launches}, min resid {out['resid'].min():.3e}",
          flush=True)
bd = eng.profile_breakdown()
for k, v in bd.items():
    if v['launches']:
        extra = f"  {v['work'] / v['ms'] / 1e9:.2f} TFLOP/s" if k in ('lu_gemm', 'matvec_gemm', 'panel') and v['ms'] > 0 else ''
        print(f"  {k:12s} {v['ms'] / steps:9.3f} ms/step  launches/step {v['launches'] // steps}{extra}")
eng.close()

"""Per-phase clock64() sums inside lu_panel_kernel (instrumented build: `make -C adaptive-matrix-solver_b200 PROF=1`).
Usage: python profiles/panel_phases.py [C] [n]      (diagnostic only, never a bench number)"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_matrix_solver_b200 import _abi                   # noqa: E402
_abi.LIB_NAME = "libmaus_b200_prof.so"
import adaptive_matrix_solver_b200 as pkg                      # noqa: E402
from adaptive_matrix_solver_b200.workloads import k2_matrix, initial_vectors   # noqa: E402

C_ = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
A = k2_matrix(n)
V = initial_vectors(C_, n)
eng = pkg.MausEngine(0)
lib = _abi.load_library()
lib.maus_debug_panel_prof.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
eng.set_matrix(A)
eng.upload_vectors(V)
alpha = np.full(C_, 0.01); psi = np.full(C_, 1e-20)
eng.step(_abi.EIGENVALUE, alpha, psi, V=None, rng_key=np.arange(C_, dtype=np.uint64))
buf = (ctypes.c_uint64 * 16)()
lib.maus_debug_panel_prof(buf, 1)
eng.profile_reset(True)
eng.step(_abi.EIGENVALUE, alpha, psi, V=None, rng_key=np.arange(C_, dtype=np.uint64) + np.uint64(1000))
bd = eng.profile_breakdown()
lib.maus_debug_panel_prof(buf, 0)
names = ["block start / loads issued", "local + warp search + CTA barrier (first column: waits for the loads)", "CTA reduce + publish (DSMEM)",
         "cluster barrier (column)", "read slots + apply column", "cluster barrier + L11 gather", "U12 = L11^-1 rows",
         "cluster barrier (U12)", "U12 write-back + rank-IB update"]
tot = sum(buf[i] for i in range(9))
print(f"C={C_} n={n}: panel {bd['panel']['ms']:.2f} ms per step over {bd['panel']['launches']} launches; instrumented thread: {buf[15]} launches, "
      f"{tot / 1.965e6:.2f} ms of clocks at 1965 MHz")
for i, nm in enumerate(names):
    print(f"  {i}: {buf[i] / 1.965e6:8.3f} ms  {100.0 * buf[i] / max(tot, 1):5.1f} %  {nm}")
eng.close()

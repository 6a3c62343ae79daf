import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import adaptive_matrix_solver_b200 as pkg
from adaptive_matrix_solver_b200 import _abi
from adaptive_matrix_solver_b200.workloads import k5_sparse
n, C_ = 1_000_000, 8
A = k5_sparse(n)
eng = pkg.MausEngine(0)
eng.set_matrix(A)
rng = np.random.default_rng(1)
V = rng.random((C_, n)) + 1j * rng.random((C_, n)); V /= np.linalg.norm(V, axis=1, keepdims=True)
eng.upload_vectors(V)
X, st, it = eng.solve_shifted(np.zeros(C_, dtype=complex), np.full(C_, 5e-19), rng_key=None, method=_abi.METHOD_GMRES, RHS=None)
print(st, it)
eng.close()

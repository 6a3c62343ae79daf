"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck), sizes kept tiny."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaptive_matrix_solver_b200 as pkg                      # noqa: E402
from adaptive_matrix_solver_b200 import _abi                   # noqa: E402
from adaptive_matrix_solver_b200.workloads import k2_matrix, initial_vectors, k5_sparse   # noqa: E402

eng = pkg.MausEngine(0)
rng = np.random.default_rng(0)
for n, C_ in ((300, 3), (200, 13), (129, 2)):          # panel modes A / B, partial last panel, outer-block groups
    A = k2_matrix(n, seed=n)
    eng.set_matrix(A)
    V = initial_vectors(C_, n, seed=n)
    out = eng.step(_abi.EIGENVALUE, np.full(C_, 0.3), np.full(C_, 1e-20), V=V, rng_key=np.arange(C_, dtype=np.uint64))
    assert (out["status"] == 0).all() and np.isfinite(out["resid"]).all()
    X, st, it = eng.solve_shifted(out["lam"], np.full(C_, 1e-19), rng_key=None, method=_abi.METHOD_GMRES, use_jacobi=np.ones(C_, np.uint8), RHS=V)
    print("dense", n, C_, out["resid"].max(), st.tolist(), it.tolist(), flush=True)
As = k5_sparse(500, nnz_per_row=6, seed=2)
eng.set_matrix(As)
b = rng.standard_normal((3, 500)) + 0j
X, st, it = eng.solve_shifted(np.zeros(3, complex), np.full(3, 1e-19), rng_key=None, method=_abi.METHOD_GMRES, RHS=b)
print("sparse", st.tolist(), it.tolist(), flush=True)
Ar = rng.standard_normal((70, 45)) + 1j * rng.standard_normal((70, 45))
eng.svd_set_matrix(Ar)
U = np.ascontiguousarray(rng.standard_normal((10, 70)) + 0j); Vv = np.ascontiguousarray(rng.standard_normal((10, 45)) + 0j)
o = eng.svd_step(U, Vv)
print("svd", o["status"].tolist(), float(o["resid"].max()), flush=True)
eng.close()
print("done")

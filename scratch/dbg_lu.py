import numpy as np, sys
sys.path.insert(0, '/root/repo')
import adaptive_matrix_solver_b200 as pkg
eng = pkg.MausEngine(0)
rng = np.random.default_rng(0)
for n in (1, 2, 5, 8, 16, 17, 40, 100, 128, 129, 256, 300):
    A = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) + 3*np.eye(n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    eng.set_matrix(A)
    X, st, _ = eng.solve_shifted([0j], [0.0], rng_key=None, RHS=b[None, :])
    xr = np.linalg.solve(A, b)
    err = np.linalg.norm(X[0] - xr) / np.linalg.norm(xr)
    res = np.linalg.norm(A @ X[0] - b) / np.linalg.norm(b)
    print(f"n={n:4d} status={st[0]} relerr={err:.3e} resid={res:.3e}", flush=True)

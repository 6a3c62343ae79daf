"""Synthetic problem families of the benchmark configurations (BASELINE.json configs)."""
import numpy as np


def k2_matrix(n, seed=20260):
    """K2 / K3: A = G/sqrt(n) + D, G iid complex with re, im ~ U(-1/2, 1/2), D = diag(linspace(-2,2) + i linspace(-1,1)):
    dense non-Hermitian, well separated spectrum, cond ~ 1e1-1e2 ('Stable' -> direct solve)."""
    rng = np.random.default_rng(seed)
    G = (rng.random((n, n)) - 0.5) + 1j * (rng.random((n, n)) - 0.5)
    A = G / np.sqrt(n)
    A[np.arange(n), np.arange(n)] += np.linspace(-2, 2, n) + 1j * np.linspace(-1, 1, n)
    return A


def initial_vectors(C, n, seed=20260):
    """Unit random start vectors drawn like AMS:130-134 (rand + i rand, normalised), from a private generator."""
    rng = np.random.default_rng(seed + 7)
    V = rng.random((C, n)) + 1j * rng.random((C, n))
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    return np.ascontiguousarray(V, dtype=np.complex128)


def k4_system(n, seed=20260):
    """K4: ill-conditioned dense Ax=b, A = Q1 diag(logspace(0,-9)) Q2^H + 1e-3 D with Householder Q's (cond ~ 1e9 ->
    'Fragile' -> GMRES preferred), b = A 1."""
    rng = np.random.default_rng(seed)
    u = rng.standard_normal(n) + 1j * rng.standard_normal(n); u /= np.linalg.norm(u)
    w = rng.standard_normal(n) + 1j * rng.standard_normal(n); w /= np.linalg.norm(w)
    s = np.logspace(0, -9, n)
    # Q1 S Q2^H with Q = I - 2 u u^H, applied without forming Q
    S = np.diag(s).astype(np.complex128)
    Q1S = S - 2.0 * np.outer(u, u.conj() @ S)
    A = Q1S - 2.0 * np.outer(Q1S @ w, w.conj())
    A[np.arange(n), np.arange(n)] += 1e-3 * (np.linspace(1, 2, n) + 1j * np.linspace(-1, 1, n))
    b = A @ np.ones(n, dtype=np.complex128)
    return A, b


def k5_sparse(n, nnz_per_row=20, seed=20260):
    """K5: CSC, ~nnz_per_row entries per row at uniform random columns, values re, im ~ U(-1/2, 1/2), plus a dominant
    diagonal 4 + linspace so GMRES(20) converges in about one cycle."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n, dtype=np.int64), nnz_per_row)
    cols = rng.integers(0, n, size=n * nnz_per_row, dtype=np.int64)
    vals = (rng.random(n * nnz_per_row) - 0.5) + 1j * (rng.random(n * nnz_per_row) - 0.5)
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsc()
    A = A + sp.diags(4.0 + np.linspace(0, 1, n) + 0j, format="csc")
    return sp.csc_matrix(A)

"""Set-up diagnostics on the device: maus_diag_dense / maus_cond2_estimate behind diagnostics.diagnose_matrix_initial
against numpy's own np.allclose / np.count_nonzero / np.linalg.cond (what AMS:374-404 calls), and the three-way strategy
decision of AMS:405-421 on the K2 / K4 matrix families."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import adaptive_matrix_solver_b200 as pkg
    e = pkg.MausEngine(0)
    yield e
    e.close()


def _reference_diag(M):
    """AMS:374-404 restated with the numpy calls the reference makes (dense ndarray branch)"""
    d = {'is_hermitian': False, 'is_complex_symmetric': False, 'is_sparse_init': False, 'condition_number': np.inf,
         'is_singular': False}
    d['is_sparse_init'] = (np.count_nonzero(M) / M.size) < 0.25
    if np.allclose(M, M.conj().T):
        d['is_hermitian'] = True
    if np.allclose(M, M.T):
        d['is_complex_symmetric'] = True
    if not d['is_sparse_init']:
        c = np.linalg.cond(M)
        d['condition_number'] = c
        d['is_singular'] = bool(np.isinf(c) or c > 1e15)
    return d


def _cases():
    from adaptive_matrix_solver_b200.workloads import k2_matrix, k4_system
    rng = np.random.default_rng(0)
    out = []
    out.append(("k2-256", k2_matrix(256, seed=1)))
    out.append(("k2-1024", k2_matrix(1024, seed=2)))
    out.append(("k4-512", k4_system(512)[0]))                    # cond ~ 1e6..1e9 -> Fragile
    G = rng.standard_normal((300, 300)) + 1j * rng.standard_normal((300, 300))
    out.append(("hermitian", G + G.conj().T))
    out.append(("complex-symmetric", G + G.T))
    out.append(("real-symmetric", (G + G.T).real.astype(np.complex128)))
    U, _, Vh = np.linalg.svd(G)
    out.append(("cond-1e13", (U * np.logspace(0, -13, 300)) @ Vh))  # Critical
    S = G.copy(); S[:, 7] = S[:, 3]
    out.append(("singular", S))
    D = np.zeros((400, 400), dtype=np.complex128); D[np.arange(400), np.arange(400)] = np.linspace(1, 2, 400)
    out.append(("mostly-zero", D))                               # is_sparse_init -> no condition number
    H = G + G.conj().T; H[5, 9] += 1e-3                          # Hermitian up to one entry beyond the allclose tolerance
    out.append(("almost-hermitian", H))
    return out


@pytest.mark.parametrize("name,M", _cases(), ids=[c[0] for c in _cases()])
def test_device_diagnosis_reproduces_the_reference_decisions(eng, name, M):
    from adaptive_matrix_solver_b200.diagnostics import diagnose_matrix_initial, initial_strategy
    ref = _reference_diag(M)
    got = diagnose_matrix_initial(eng, M)
    for k in ("is_hermitian", "is_complex_symmetric", "is_sparse_init", "is_singular"):
        assert got[k] == ref[k], (name, k, got, ref)
    if np.isfinite(ref["condition_number"]) and ref["condition_number"] < 1e15:
        # one-sided estimate: never above the true 2-norm condition number (up to rounding), tight to a few per cent
        assert got["condition_number"] <= ref["condition_number"] * (1 + 1e-2), (name, got, ref)
        assert got["condition_number"] >= 0.9 * ref["condition_number"], (name, got["condition_number"], ref["condition_number"])
    for ptype in ("EIGENVALUE", "SOLVE_LINEAR_SYSTEM", "SVD"):
        assert initial_strategy(got, ptype, 1e-8) == initial_strategy(ref, ptype, 1e-8), (name, ptype)


def test_condition_estimate_at_full_order(eng):
    """n = 2048 with a known, densely clustered singular spectrum (neighbours 0.5 % apart): the estimate costs a few batched LU
    solves instead of the SVD and still lands within a few per cent of both ends."""
    n = 2048
    rng = np.random.default_rng(5)
    # Householder-product orthogonal factors applied to a known singular spectrum (cond = 3.7e4)
    s = np.geomspace(37.0, 1e-3, n)
    u = rng.standard_normal(n) + 1j * rng.standard_normal(n); u /= np.linalg.norm(u)
    w = rng.standard_normal(n) + 1j * rng.standard_normal(n); w /= np.linalg.norm(w)
    S = np.diag(s).astype(np.complex128)
    A = S - 2.0 * np.outer(u, u.conj() @ S)
    A = A - 2.0 * np.outer(A @ w, w.conj())
    eng.set_matrix(A)
    smax, smin, st = eng.cond2_estimate()
    assert st == 0
    assert 0.95 * 37.0 <= smax <= 37.0 * (1 + 1e-9) and 1e-3 * (1 - 1e-6) <= smin <= 1.05e-3

"""BASELINE.json configurations 4 and 5 at FULL size through the C ABI: size-independent properties (the oracle would need
minutes per candidate here) plus one scipy cross-check where it is cheap.  K2 / K3 live in test_gpu_step_parity.py /
test_gpu_edges.py."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import adaptive_matrix_solver_b200 as pkg
    e = pkg.MausEngine(0)
    yield e
    e.close()


@pytest.mark.timeout(600)
def test_k4_dense_ill_conditioned_gmres_with_jacobi_full_size(eng):
    """Config 4: dense Ax = b, n = 8192, cond ~ 1e9, GMRES(20) x 50 with the Jacobi preconditioner for stuck candidates
    (AMS:61-90).  Every candidate must meet scipy's stopping rule, Jacobi must cut the iteration count, and the
    preconditioned solve must replay scipy's iteration count and solution."""
    from adaptive_matrix_solver_b200 import _abi
    from adaptive_matrix_solver_b200.workloads import k4_system
    n, C = 8192, 12
    A, b = k4_system(n)
    eng.set_matrix(A); eng.set_rhs(b)
    rng = np.random.default_rng(0)
    X0 = rng.standard_normal((C, n)) + 1j * rng.standard_normal((C, n))
    eng.upload_vectors(X0)
    jac = (np.arange(C) % 2).astype(np.uint8)
    psi = np.full(C, 1e-19)
    X, st, it = eng.solve_shifted(np.zeros(C, dtype=complex), psi, rng_key=np.arange(C, dtype=np.uint64), method=_abi.METHOD_GMRES,
                                  use_jacobi=jac, RHS=None, rhs_shared=True)
    assert (st == 0).all()
    nb = np.linalg.norm(b)
    for c in range(C):
        assert np.linalg.norm(A @ X[c] - b) <= 1e-8 * nb * (1 + 1e-6)
    assert it[jac == 1].max() < it[jac == 0].min()              # the preconditioner really acts
    assert len(set(it[jac == 1].tolist())) == 1 and len(set(it[jac == 0].tolist())) == 1      # same system, same count
    # scipy replay of one Jacobi candidate (x0 = b like AMS:89; a handful of dense matvecs)
    d = np.diag(A) + psi[1]
    M = spla.LinearOperator((n, n), matvec=lambda v: v / d, dtype=np.complex128)
    cnt = []
    H = spla.LinearOperator((n, n), matvec=lambda v: A @ v + psi[1] * v, dtype=np.complex128)      # H = A + psi I (AMS:47-52)
    xr, info = spla.gmres(H, b, x0=b, rtol=1e-8, maxiter=50, M=M,
                          callback=lambda r: cnt.append(r), callback_type="pr_norm")
    assert info == 0 and len(cnt) == it[1]
    assert np.linalg.norm(X[1] - xr) <= 1e-7 * np.linalg.norm(xr)


@pytest.mark.timeout(900)
def test_k5_sparse_million_row_gmres_full_size(eng):
    """Config 5: sparse CSC, n = 1 000 000, ~21 nnz per row.  SpMM against scipy, GMRES meets rtol on every candidate with
    one iteration count (same operator), and the long-vector kernels (RQ, residual) agree with numpy."""
    from adaptive_matrix_solver_b200 import _abi
    from adaptive_matrix_solver_b200.workloads import k5_sparse
    n, C = 1_000_000, 5                                  # 5: one packed pass of 4 candidates + a single-candidate pass
    A = k5_sparse(n)
    rng = np.random.default_rng(1)
    V = rng.random((C, n)) + 1j * rng.random((C, n)); V /= np.linalg.norm(V, axis=1, keepdims=True)
    eng.set_matrix(A)
    eng.upload_vectors(V)
    lam, vn2 = eng.rq(C_=C)                                # SpMM (packed + unpacked kernels) + multi-block dots
    AV = (A @ V.T).T
    for c in range(C):
        ref = np.vdot(V[c], AV[c])
        assert abs(lam[c] - ref) <= 1e-12 * abs(ref) and abs(vn2[c] - 1) <= 1e-12
    r = eng.residual(_abi.EIGENVALUE, lam=lam, C_=C)
    for c in range(C):
        ref = np.linalg.norm(AV[c] - lam[c] * V[c])
        assert abs(r[c] - ref) <= 1e-11 * ref
    X, st, it = eng.solve_shifted(np.zeros(C, dtype=complex), np.full(C, 5e-19), rng_key=None, method=_abi.METHOD_GMRES, RHS=None)
    assert (st == 0).all() and len(set(it.tolist())) == 1 and 1 <= it[0] <= 40
    for c in range(C):
        assert np.linalg.norm(A @ X[c] - V[c]) <= 1e-8 * (1 + 1e-6)          # ||b|| = 1
    # scipy replay of candidate 0 on the same operator H = A + psi I (AMS:47-52), x0 = b (AMS:61, 89): the device GMRES must
    # take exactly scipy's number of inner iterations and land on its solution
    cnt = []
    H = (A + 5e-19 * sp.identity(n, dtype=np.complex128, format="csc")).tocsr()
    xr, info = spla.gmres(H, V[0], x0=V[0], rtol=1e-8, maxiter=50, callback=lambda r: cnt.append(r), callback_type="pr_norm")
    assert info == 0 and len(cnt) == it[0], (len(cnt), it[0])
    assert np.linalg.norm(X[0] - xr) <= 1e-7 * np.linalg.norm(xr)


@pytest.mark.timeout(600)
def test_direct_solve_at_the_maximum_order(eng):
    """n = 8192 is the largest order of the batched LU (one panel cluster owns all rows, two rows per thread): backward error
    of the shifted solves and agreement of the Rayleigh quotients with numpy."""
    from adaptive_matrix_solver_b200 import _abi
    from adaptive_matrix_solver_b200.workloads import k2_matrix, initial_vectors
    n, C = 8192, 2
    A = k2_matrix(n, seed=3)
    V = initial_vectors(C, n, seed=5)
    eng.set_matrix(A)
    eng.upload_vectors(V)
    lam, _ = eng.rq(C_=C)
    for c in range(C):
        assert abs(lam[c] - np.vdot(V[c], A @ V[c])) <= 1e-12 * abs(lam[c])
    psi = np.full(C, 1e-20)
    X, st, _ = eng.solve_shifted(lam, psi, rng_key=np.arange(C, dtype=np.uint64) + 1, method=_abi.METHOD_LU, RHS=V)
    assert (st == 0).all()
    anorm = np.abs(A).sum(axis=1).max()
    for c in range(C):
        r = A @ X[c] - (lam[c] - psi[c]) * X[c] - V[c]
        be = np.linalg.norm(r) / (anorm * np.linalg.norm(X[c]) + np.linalg.norm(V[c]))
        assert be < 1e-13, be


@pytest.mark.timeout(600)
def test_k4_population_step_through_seam_b_full_size(eng):
    """Config 4 through the drop-in boundary (step_population on candidate objects, GMRES preferred, Jacobi for the stuck
    half, AMS:61-90, 284-285, 298-299): x <- (1 - a) x + a x_solve with x_solve meeting scipy's rtol, residual = ||A x - b||."""
    import random
    from adaptive_matrix_solver_b200 import step_population
    from adaptive_matrix_solver_b200.workloads import k4_system
    from mock_candidate import MockCandidate, ProblemType
    n, C = 8192, 8
    A, b = k4_system(n)
    np.random.seed(4); random.seed(4)
    cands = [MockCandidate(A, ProblemType.SOLVE_LINEAR_SYSTEM, n) for _ in range(C)]
    for k, c in enumerate(cands):
        c.stuck_counter = 2 if k % 2 else 0                      # > 1 switches the Jacobi preconditioner on (AMS:64)
        c.alpha_local_step = 0.25
    x_before = [c.x_k.copy() for c in cands]
    strat = dict(overall_psi_aggression_factor=10.0, max_psi_retries=25, current_convergence_threshold=1e-4)
    know = dict(local_solver_preference="iterative_gmres", is_sparse_problem=False, is_hermitian=False)
    assert step_population(cands, A, b, strat, know, eng) == C
    nb = np.linalg.norm(b)
    for c, x0 in zip(cands, x_before):
        assert c.local_psi_retries_needed == 0                   # first attempt succeeded: GMRES converged for every candidate
        xs = (c.x_k - 0.75 * x0) / 0.25                          # the solve result behind the damped mix (AMS:285)
        assert np.linalg.norm(A @ xs - b) <= 2e-8 * nb
        r = np.linalg.norm(A @ c.x_k - b)
        assert abs(c.residual_k - r) <= 1e-10 * max(r, 1.0)
    assert [c.stuck_counter for c in cands] == [0, 1] * (C // 2)  # max(0, stuck - 1) on success (AMS:286)


@pytest.mark.timeout(900)
def test_k5_sparse_population_step_through_seam_b_full_size(eng):
    """Config 5 through the drop-in boundary: sparse CSC linear system, n = 1 000 000, GMRES / SpMV path, long-vector mix and
    residual kernels (AMS:46-47, 61-90, 285, 299)."""
    import random
    from adaptive_matrix_solver_b200 import step_population
    from adaptive_matrix_solver_b200.workloads import k5_sparse
    from mock_candidate import MockCandidate, ProblemType
    n, C = 1_000_000, 3
    A = k5_sparse(n)
    rng = np.random.default_rng(8)
    b = rng.random(n) + 1j * rng.random(n)
    np.random.seed(6); random.seed(6)
    cands = [MockCandidate(A, ProblemType.SOLVE_LINEAR_SYSTEM, n) for _ in range(C)]
    for c in cands:
        c.alpha_local_step = 0.5
    x_before = [c.x_k.copy() for c in cands]
    strat = dict(overall_psi_aggression_factor=1.0, max_psi_retries=25, current_convergence_threshold=1e-6)
    know = dict(local_solver_preference="iterative_gmres", is_sparse_problem=True, is_hermitian=False)
    assert step_population(cands, A, b, strat, know, eng) == C
    nb = np.linalg.norm(b)
    for c, x0 in zip(cands, x_before):
        assert c.local_psi_retries_needed == 0
        xs = 2.0 * c.x_k - x0                                    # the solve result behind the damped mix (alpha = 0.5)
        assert np.linalg.norm(A @ xs - b) <= 2e-8 * nb
        r = np.linalg.norm(A @ c.x_k - b)
        assert abs(c.residual_k - r) <= 1e-10 * max(r, 1.0)


@pytest.mark.parametrize("C", [2, 3, 8, 9, 13, 16])
def test_spmm_candidate_groups_match_scipy(eng, C):
    """SpMM through every grouping of the candidate block (groups of 8 through the 128-byte-per-entry layout, groups of 4 / 2,
    a single candidate behind them) on a ragged matrix: empty rows, one-entry rows, rows longer than one 24-entry chunk.  The
    Rayleigh quotient and the residual norm see every entry of A V; n is above the multi-block threshold of the reductions."""
    from adaptive_matrix_solver_b200 import _abi
    n = 40_000
    rng = np.random.default_rng(11 + C)
    counts = rng.choice([0, 1, 5, 21, 24, 25, 49, 60], size=n, p=[0.05, 0.05, 0.2, 0.4, 0.1, 0.1, 0.05, 0.05])
    rows = np.repeat(np.arange(n), counts)
    cols = rng.integers(0, n, size=rows.size)
    vals = (rng.random(rows.size) - 0.5) + 1j * (rng.random(rows.size) - 0.5)
    A = sp.csc_matrix(sp.coo_matrix((vals, (rows, cols)), shape=(n, n)))
    V = rng.random((C, n)) + 1j * rng.random((C, n)); V /= np.linalg.norm(V, axis=1, keepdims=True)
    eng.set_matrix(A)
    eng.upload_vectors(V)
    lam, vn2 = eng.rq(C_=C)
    AV = (A @ V.T).T
    for c in range(C):
        ref = np.vdot(V[c], AV[c])
        assert abs(lam[c] - ref) <= 1e-12 * max(abs(ref), 1.0), (c, lam[c], ref)
        assert abs(vn2[c] - 1) <= 1e-12
    r = eng.residual(_abi.EIGENVALUE, lam=lam, C_=C)
    for c in range(C):
        ref = np.linalg.norm(AV[c] - lam[c] * V[c])
        assert abs(r[c] - ref) <= 1e-11 * ref, (c, r[c], ref)

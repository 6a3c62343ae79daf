"""GPU parity of the batched GMRES path (AMS:61-90 -> scipy.sparse.linalg.gmres) against scipy through the oracle."""
import random
import warnings

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from golden_io import Golden
from mock_candidate import MockCandidate, ProblemType
from oracle import maus_oracle as mo
from parity import anorm, assert_scalar_close, vec_err_up_to_phase

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import adaptive_matrix_solver_b200 as pkg
    e = pkg.MausEngine(0)
    yield e
    e.close()


def crand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def scipy_gmres(H, b, M=None):
    counter = []
    x, info = spla.gmres(H, b, x0=b, rtol=1e-8, maxiter=50, M=M, callback=lambda r: counter.append(r),
                         callback_type="pr_norm")
    return x, info, len(counter)


@pytest.mark.parametrize("n,C", [(5, 2), (64, 3), (300, 9), (1024, 16)])
def test_dense_gmres_matches_scipy(eng, n, C):
    from adaptive_matrix_solver_b200 import _abi
    rng = np.random.default_rng(n)
    A = crand(rng, n, n) / np.sqrt(n) + 3.0 * np.eye(n)
    RHS = crand(rng, C, n)
    sigma = 0.3 * crand(rng, C)
    psi = np.full(C, 1e-19)
    eng.set_matrix(A)
    X, st, it = eng.solve_shifted(sigma, psi, rng_key=np.arange(C) + 1, method=_abi.METHOD_GMRES, RHS=RHS)
    for c in range(C):
        H = A - sigma[c] * np.eye(n) + psi[c] * np.eye(n)
        xr, info, nit = scipy_gmres(H, RHS[c])
        assert (st[c] == 0) == (info == 0)
        assert it[c] == nit, (c, it[c], nit)
        assert np.linalg.norm(X[c] - xr) <= 1e-10 * np.linalg.norm(xr)
        assert np.linalg.norm(H @ X[c] - RHS[c]) <= 1e-8 * np.linalg.norm(RHS[c]) * (1 + 1e-6)


def test_jacobi_preconditioned_gmres_matches_scipy(eng):
    """AMS:64-86: M = diag(1/diag(H)) when stuck_counter > 1 and the diagonal is finite and > 1e-12."""
    from adaptive_matrix_solver_b200 import _abi
    n, C = 200, 4
    rng = np.random.default_rng(11)
    d = np.logspace(0, 4, n) * np.exp(1j * rng.uniform(0, 0.3, n))
    A = np.diag(d) + 0.05 * crand(rng, n, n)
    RHS = crand(rng, C, n)
    eng.set_matrix(A)
    jac = np.array([1, 0, 1, 0], dtype=np.uint8)
    X, st, it = eng.solve_shifted(np.zeros(C, dtype=complex), np.full(C, 1e-19), rng_key=None, method=_abi.METHOD_GMRES,
                                  use_jacobi=jac, RHS=RHS)
    for c in range(C):
        H = A + 1e-19 * np.eye(n)
        M = np.diag(1.0 / np.diag(H)) if jac[c] else None
        xr, info, nit = scipy_gmres(H, RHS[c], M)
        assert (st[c] == 0) == (info == 0), (c, st[c], info)
        assert it[c] == nit, (c, it[c], nit)
        if info == 0:
            assert np.linalg.norm(X[c] - xr) <= 1e-9 * np.linalg.norm(xr)
    assert it[0] < it[1]            # the preconditioner really changes the iteration


def test_jacobi_rejected_for_tiny_diagonal(eng):
    """|d| <= 1e-12 anywhere -> preconditioner is dropped (AMS:72), same iteration as the unpreconditioned solve."""
    from adaptive_matrix_solver_b200 import _abi
    n = 60
    rng = np.random.default_rng(2)
    A = crand(rng, n, n) / np.sqrt(n) + 3 * np.eye(n)
    A[7, 7] = 1e-14
    b = crand(rng, 1, n)
    eng.set_matrix(A)
    X1, st1, it1 = eng.solve_shifted([0j], [0.0], rng_key=None, method=_abi.METHOD_GMRES, use_jacobi=[1], RHS=b)
    X2, st2, it2 = eng.solve_shifted([0j], [0.0], rng_key=None, method=_abi.METHOD_GMRES, use_jacobi=[0], RHS=b)
    assert it1[0] == it2[0] and np.array_equal(X1, X2)


def test_sparse_gmres_matches_scipy(eng):
    from adaptive_matrix_solver_b200 import _abi
    from adaptive_matrix_solver_b200.workloads import k5_sparse
    n, C = 3000, 6
    A = k5_sparse(n, seed=5)
    rng = np.random.default_rng(3)
    RHS = crand(rng, C, n)
    sigma = 0.2 * crand(rng, C)
    psi = np.full(C, 5e-19)
    eng.set_matrix(A)
    X, st, it = eng.solve_shifted(sigma, psi, rng_key=None, method=_abi.METHOD_GMRES, RHS=RHS)
    for c in range(C):
        H = sp.csc_matrix(A - sigma[c] * sp.eye(n, format="csc") + psi[c] * sp.identity(n, format="csc"))
        xr, info, nit = scipy_gmres(H, RHS[c])
        assert st[c] == 0 and info == 0
        assert it[c] == nit
        assert np.linalg.norm(X[c] - xr) <= 1e-10 * np.linalg.norm(xr)


def test_gmres_nonconvergence_reports_status(eng):
    """A rotation-like operator GMRES(20) x 50 cannot solve to 1e-8 -> info != 0 -> LinAlgError in the reference (AMS:90)."""
    from adaptive_matrix_solver_b200 import _abi
    n = 400
    A = np.roll(np.eye(n), 1, axis=1).astype(np.complex128)      # cyclic shift: GMRES stagnates for n-1 steps
    b = np.zeros((1, n), dtype=np.complex128); b[0, 0] = 1.0
    eng.set_matrix(A)
    X, st, it = eng.solve_shifted([0j], [0.0], rng_key=None, method=_abi.METHOD_GMRES, RHS=b)
    xr, info, nit = scipy_gmres(A, b[0])
    assert info != 0 and st[0] == _abi.ST_GMRES_NOCONV
    assert it[0] == nit


@pytest.mark.parametrize("name,stride", [("gmres64", 1), ("speig200", 1), ("lin5_shim", 1)])
def test_step_population_replays_reference_gmres_golden(eng, name, stride):
    """Reference steps recorded with the tol->rtol shim (GMRES really runs, Jacobi for stuck > 1, sparse Psi)."""
    from adaptive_matrix_solver_b200 import step_population
    g = Golden(name)
    ptype = ProblemType(g.problem_type)
    floor = 4e-13 * max(anorm(g.A), 1.0)
    checked = 0
    for i in range(0, g.n_steps, stride):
        before, after = g.side("before", i), g.side("after", i)
        c = MockCandidate.__new__(MockCandidate)
        c.id = int(g.z["cand_id"][i]); c.N_diag = g.n; c.problem_type = ptype
        c.problem_matrix = g.ctor_matrix(i); c.b_vector = None
        c.load(before)
        seed = int(g.z["seed"][i]); np.random.seed(seed % 2 ** 32); random.seed(seed)
        step_population([c], g.A, g.b, g.strat(i), g.know(i), eng)
        assert c.local_psi_retries_needed == after["retries"], (name, i)
        assert c.stuck_counter == after["stuck"], (name, i)
        # GMRES answers are only rtol=1e-8 accurate: two correct implementations agree to about
        # cond * eps on x; compare what the candidate step consumes with that floor
        gfloor = max(floor, 1e-7 * max(1.0, abs(after["res"])))
        if g.problem_type == 1:
            assert_scalar_close(c.lambda_k, after["lam"], floor, f"{name}[{i}] lambda")
            assert vec_err_up_to_phase(c.v_k, after["v"]) <= 1e-7, (name, i)
        else:
            assert np.abs(c.x_k - after["x"]).max() <= 1e-7 * np.abs(after["x"]).max() + floor, (name, i)
        assert abs(c.residual_k - after["res"]) <= 1e-6 * abs(after["res"]) + gfloor, (name, i, c.residual_k, after["res"])
        checked += 1
    assert checked > 20


def test_matvec_compaction_keeps_every_candidate_on_scipys_iteration(eng):
    """Candidates that finish early (Jacobi) are dropped from the batched matvec (gmres.cu: matvec compaction); the ones still
    iterating must not notice: same iteration counts as scipy, same solutions as when they are solved without fast peers."""
    from adaptive_matrix_solver_b200 import _abi
    n, C = 320, 16
    rng = np.random.default_rng(31)
    d = np.logspace(0, 3.5, n) * np.exp(1j * rng.uniform(0, 0.3, n))
    A = np.diag(d) + 0.05 * crand(rng, n, n)
    RHS = crand(rng, C, n)
    jac = (np.arange(C) % 4 != 0).astype(np.uint8)               # 12 fast (Jacobi) candidates, 4 slow ones
    eng.set_matrix(A)
    zero, psi = np.zeros(C, dtype=complex), np.full(C, 1e-19)
    X, st, it = eng.solve_shifted(zero, psi, rng_key=None, method=_abi.METHOD_GMRES, use_jacobi=jac, RHS=RHS)
    slow = np.flatnonzero(jac == 0)
    assert it[jac == 1].max() < 20 < it[slow].min()               # the slow ones run several restart cycles alone
    Xs, sts, its = eng.solve_shifted(zero[slow], psi[slow], rng_key=None, method=_abi.METHOD_GMRES, use_jacobi=jac[slow], RHS=RHS[slow])
    assert np.array_equal(it[slow], its) and np.array_equal(st[slow], sts)
    for k, c in enumerate(slow):
        assert np.linalg.norm(X[c] - Xs[k]) <= 1e-11 * np.linalg.norm(Xs[k])
    H = A + 1e-19 * np.eye(n)
    for c in (0, 1):
        M = np.diag(1.0 / np.diag(H)) if jac[c] else None
        xr, info, nit = scipy_gmres(H, RHS[c], M)
        assert (st[c] == 0) == (info == 0) and it[c] == nit

"""GPU parity tests of the individual CUDA kernels, called through the C ABI, against numpy / the oracle."""
import numpy as np
import pytest

from oracle import maus_oracle as mo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import adaptive_matrix_solver_b200 as pkg
    e = pkg.MausEngine(0)
    yield e
    e.close()


def crand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@pytest.mark.parametrize("M,N,K,batch", [(128, 64, 16, 1), (128, 64, 128, 2), (8, 8, 4, 1), (1, 1, 1, 1), (5, 3, 7, 3),
                                         (300, 129, 130, 2), (257, 65, 33, 1), (64, 200, 128, 1), (1000, 70, 1, 1)])
@pytest.mark.parametrize("beta,negate", [(0, False), (1, True)])
def test_zgemm_dmma_matches_numpy(eng, M, N, K, batch, beta, negate):
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    A = crand(rng, batch, M, K); B = crand(rng, batch, K, N); C0 = crand(rng, batch, M, N)
    ref = (C0 if beta else 0) + (-1 if negate else 1) * (A @ B)
    # the ABI takes column-major matrices: pass the transposes' memory
    Acm = np.ascontiguousarray(A.transpose(0, 2, 1)); Bcm = np.ascontiguousarray(B.transpose(0, 2, 1))
    Ccm = np.ascontiguousarray(C0.transpose(0, 2, 1))
    out = eng.debug_zgemm(Acm, Bcm, Ccm, beta=beta, negate=negate, use_dmma=True).transpose(0, 2, 1)
    out2 = eng.debug_zgemm(Acm, Bcm, Ccm, beta=beta, negate=negate, use_dmma=False).transpose(0, 2, 1)
    scale = np.abs(A).max() * np.abs(B).max() * K + 1
    assert np.abs(out - ref).max() <= 1e-14 * scale
    assert np.abs(out2 - ref).max() <= 1e-14 * scale
    # the LU's 3M kernel (three real products): normwise the same bound, a different rounding pattern
    out3 = eng.debug_zgemm(Acm, Bcm, Ccm, beta=beta, negate=negate, use_dmma=2).transpose(0, 2, 1)
    assert np.abs(out3 - ref).max() <= 2e-14 * scale
    # ... and its 128 x 32 tile shape (skinny batched A*V); same products, same accumulation order per element
    out4 = eng.debug_zgemm(Acm, Bcm, Ccm, beta=beta, negate=negate, use_dmma=3).transpose(0, 2, 1)
    assert np.array_equal(out4, out3)


@pytest.mark.parametrize("n", [1, 2, 5, 8, 16, 100, 127, 128, 129, 256, 300, 520])
def test_lu_solve_with_host_R_matches_lapack(eng, n):
    """debug entry maus_solve_with_R: identical H on both sides -> x must agree to LU rounding (AMS:49-59)."""
    rng = np.random.default_rng(n)
    A = crand(rng, n, n) / np.sqrt(n) + np.diag(np.linspace(-2, 2, n) + 1j * np.linspace(-1, 1, n))
    v = crand(rng, n); v /= np.linalg.norm(v)
    lam = mo.rayleigh_quotient(A, v)
    psi = 1e-20
    np.random.seed(n)
    R = (np.random.rand(n, n) - 0.5 + 1j * (np.random.rand(n, n) - 0.5)) * psi * 0.15   # AMS:49
    eng.set_matrix(A)
    x, st = eng.solve_with_R(lam, psi, R, v)
    assert st == 0
    xr = mo.shifted_solve_dense(A, lam, psi, v, R)
    H = A - lam * np.eye(n) + psi * np.eye(n) + R
    # backward error of both solutions, and forward agreement relative to conditioning
    be = np.linalg.norm(H @ x - v) / (np.linalg.norm(H, 2) * np.linalg.norm(x) + np.linalg.norm(v))
    assert be < 1e-14 * max(4, n ** 0.5)
    cond = np.linalg.cond(H)
    assert np.linalg.norm(x - xr) <= 50 * cond * 2.2e-16 * np.linalg.norm(xr)


def test_lu_pivoting_needed(eng):
    """A matrix whose leading entries are tiny forces row exchanges in every panel column."""
    n = 200
    rng = np.random.default_rng(7)
    A = crand(rng, n, n)
    A[np.arange(n), np.arange(n)] = 1e-14          # tiny diagonal: no-pivot LU would blow up
    b = crand(rng, n)
    eng.set_matrix(A)
    X, st, _ = eng.solve_shifted([0j], [0.0], rng_key=None, RHS=b[None, :])
    assert st[0] == 0
    xr = np.linalg.solve(A, b)
    assert np.linalg.norm(X[0] - xr) <= 1e-10 * np.linalg.norm(xr)


def test_lu_zero_pivot_and_nonfinite_status(eng):
    from adaptive_matrix_solver_b200 import _abi
    n = 40
    rng = np.random.default_rng(3)
    A = crand(rng, n, n)
    A[:, 7] = 0.0                                   # exactly singular -> LAPACK info > 0 -> LinAlgError (AMS:98)
    eng.set_matrix(A)
    X, st, _ = eng.solve_shifted([0j], [0.0], rng_key=None, RHS=crand(rng, 1, n))
    assert st[0] == _abi.ST_ZERO_PIVOT
    A = crand(rng, n, n); A[3, 4] = np.nan
    eng.set_matrix(A)
    X, st, _ = eng.solve_shifted([0j], [0.0], rng_key=None, RHS=crand(rng, 1, n))
    assert st[0] == _abi.ST_NONFINITE


def test_cluster_backsolve_small_batch_statuses_and_ragged_order(eng):
    """Batches that leave SMs idle use a cluster of CTAs per candidate for the back substitution (block-cyclic rows, x blocks
    through DSMEM): ragged order (last block of 13 rows), several candidates, per-candidate status words."""
    from adaptive_matrix_solver_b200 import _abi
    n, C = 333, 3
    rng = np.random.default_rng(21)
    A = crand(rng, n, n) / np.sqrt(n) + np.diag(np.linspace(-2, 2, n) + 1j * np.linspace(-1, 1, n))
    RHS = crand(rng, C, n)
    RHS[1, 17] = np.nan                               # candidate 1 only: non-finite solution
    sigma = np.array([0.1 + 0.2j, -0.3j, 0.7], dtype=complex)
    eng.set_matrix(A)
    X, st, _ = eng.solve_shifted(sigma, np.zeros(C), rng_key=None, RHS=RHS)
    assert st[0] == 0 and st[2] == 0 and st[1] == _abi.ST_NONFINITE
    for c in (0, 2):
        xr = np.linalg.solve(A - sigma[c] * np.eye(n), RHS[c])
        assert np.linalg.norm(X[c] - xr) <= 1e-11 * np.linalg.norm(xr)
    A2 = A.copy(); A2[:, 200] = 0.0                   # exactly singular -> zero pivot reported through the cluster kernel too
    eng.set_matrix(A2)
    X, st, _ = eng.solve_shifted([0j], [0.0], rng_key=None, RHS=crand(rng, 1, n))
    assert st[0] == _abi.ST_ZERO_PIVOT


@pytest.mark.parametrize("n,C", [(8, 1), (100, 3), (256, 9), (1000, 12), (512, 64)])
def test_rq_and_residual_match_oracle(eng, n, C):
    rng = np.random.default_rng(n + C)
    A = crand(rng, n, n) / np.sqrt(n)
    V = crand(rng, C, n)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    eng.set_matrix(A)
    lam, vn2 = eng.rq(V)
    for c in range(C):
        lr = mo.rayleigh_quotient(A, V[c])
        assert abs(lam[c] - lr) <= 1e-13 * (abs(lr) + np.linalg.norm(A, 2))
        assert abs(vn2[c] - 1.0) < 1e-13
    res = eng.residual(1, V, lam)
    for c in range(C):
        rr = mo.residual_eigen(A, V[c], lam[c])
        assert abs(res[c] - rr) <= 1e-12 * rr + 1e-14


def test_psi_perturbation_is_bounded_and_keyed(eng):
    """The device Philox stream replaces np.random.rand of AMS:49: entries lie in 0.15*psi*[-0.5,0.5) and differ
    per key; checked through the solve itself with a large psi on a diagonal matrix."""
    n = 64
    A = np.diag(np.full(n, 2.0 + 0j))
    eng.set_matrix(A)
    rhs = np.ones((2, n), dtype=np.complex128)
    psi = 0.5
    X1, st, _ = eng.solve_shifted([0j, 0j], [psi, psi], rng_key=[11, 12], RHS=rhs)
    X2, _, _ = eng.solve_shifted([0j, 0j], [psi, psi], rng_key=[11, 12], RHS=rhs)
    assert np.array_equal(X1, X2)                   # deterministic for a key
    assert not np.allclose(X1[0], X1[1])            # keys decorrelate candidates
    # H = (2 + psi) I + R, |R_ij| <= 0.15*psi*0.5*sqrt(2): Neumann bound on the deviation from rhs/(2+psi)
    x0 = 1.0 / (2.0 + psi)
    assert np.abs(X1 - x0).max() < 0.15 * psi * n * x0 / (2.0 + psi)
    X3, _, _ = eng.solve_shifted([0j], [psi], rng_key=None, RHS=rhs[:1])
    assert np.abs(X3 - x0).max() < 1e-15            # no key -> no perturbation (sparse semantics, AMS:47)


@pytest.mark.parametrize("C,n", [(1, 1), (2, 7), (5, 100), (37, 4096), (130, 513)])
def test_gram_matches_numpy_vdot(eng, C, n):
    """maus_gram: G[i][j] = np.vdot(v_i, v_j) (the dedup similarity tests AMS:436, 450, 515, 520 as one device pass)."""
    rng = np.random.default_rng(C * 100 + n)
    V = crand(rng, C, n)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    G = eng.gram(V)
    ref = V.conj() @ V.T
    assert np.abs(G - ref).max() <= 1e-14
    assert np.abs(np.diag(G) - 1.0).max() <= 1e-14 and np.abs(G - G.conj().T).max() <= 1e-15


def test_device_vdot_proxy_answers_from_the_gram_matrix(eng):
    """dedup.device_vdot: vdot on converged candidates' vectors comes from the device, anything else falls through."""
    import types
    from adaptive_matrix_solver_b200.dedup import device_vdot
    from mock_candidate import MockCandidate, ProblemType
    rng = np.random.default_rng(4)
    n = 300
    A = crand(rng, n, n)
    cands = [MockCandidate(A, ProblemType.EIGENVALUE, n) for _ in range(6)]
    for k, c in enumerate(cands):
        c.state = MockCandidate.State.CONVERGED if k != 3 else MockCandidate.State.EXPLORING
    cands[4].v_k = cands[0].v_k * np.exp(0.7j)                      # a duplicate up to phase
    mod = types.SimpleNamespace(np=np)
    with device_vdot(mod, cands, eng) as px:
        assert mod.np is px
        for i in (0, 1, 2, 4, 5):
            for j in (0, 1, 2, 4, 5):
                assert abs(mod.np.vdot(cands[i].v_k, cands[j].v_k) - np.vdot(cands[i].v_k, cands[j].v_k)) <= 1e-14
        assert px.hits == 25 and px.misses == 0
        assert abs(abs(mod.np.vdot(cands[4].v_k, cands[0].v_k)) - 1.0) <= 1e-14
        w = crand(rng, n)
        assert mod.np.vdot(cands[3].v_k, w) == np.vdot(cands[3].v_k, w) and px.misses == 1     # not converged -> numpy
        assert mod.np.abs(-2.0) == 2.0                                                          # everything else forwards
    assert mod.np is np

"""Device Hermitian eigensolver (heev.cu, cyclic Jacobi) against LAPACK's eigh, and the dense Hermitian shortcut of
update_solution_step (AMS:155-186) through step_population against a restatement of those reference lines."""
import random

import numpy as np
import pytest
import scipy.linalg as sla

from mock_candidate import MockCandidate, ProblemType

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import adaptive_matrix_solver_b200 as pkg
    e = pkg.MausEngine(0)
    yield e
    e.close()


def _hermitian(n, seed, spread=1.0):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H = (G + G.conj().T) / np.sqrt(n)
    H[np.arange(n), np.arange(n)] += spread * np.linspace(-3, 3, n)
    return H


@pytest.mark.parametrize("n", [1, 2, 3, 8, 65, 300, 1024])
def test_heev_matches_lapack(eng, n):
    H = _hermitian(n, n)
    w, E = eng.heev(H)
    wr = sla.eigh(H, eigvals_only=True)
    scale = max(np.abs(wr).max(), 1e-300)
    assert np.abs(w - wr).max() <= 50 * np.finfo(float).eps * scale * max(1.0, np.sqrt(n))
    assert np.linalg.norm(H @ E - E * w) <= 1e-12 * np.linalg.norm(H) + 1e-300
    assert np.linalg.norm(E.conj().T @ E - np.eye(n)) <= 1e-12 * max(1, n)
    assert eng.heev_info["sweeps"] <= 14


def test_heev_uses_the_lower_triangle_like_eigh(eng):
    """scipy.linalg.eigh(a) reads only the lower triangle; an input that is Hermitian just to np.allclose tolerance (what
    AMS:384 accepts) must give the same spectrum as LAPACK does."""
    n = 40
    H = _hermitian(n, 7)
    noisy = H + np.triu(1e-7 * np.ones((n, n)), 1)              # perturb the strict upper triangle only
    w = eng.heev(noisy, vectors=False)
    assert np.abs(w - sla.eigh(noisy, eigvals_only=True)).max() <= 1e-13 * np.abs(w).max()
    assert np.abs(w - sla.eigh(H, eigvals_only=True)).max() <= 1e-13 * np.abs(w).max()


def test_hermitian_shortcut_population_matches_the_reference_lines(eng):
    """AMS:155-186 for a whole population in one go: ONE device eigendecomposition, the similarity scores as one GEMM, the
    residuals as one batched pass -- against the per-candidate restatement of the reference lines with LAPACK's eigh."""
    from adaptive_matrix_solver_b200 import step_population
    n, C = 96, 7
    H = _hermitian(n, 3)
    np.random.seed(4); random.seed(4)
    cands = [MockCandidate(H, ProblemType.EIGENVALUE, n) for _ in range(C)]
    v_before = [c.v_k.copy() for c in cands]
    strat = dict(overall_psi_aggression_factor=1.0, max_psi_retries=25, current_convergence_threshold=1e-10)
    know = dict(local_solver_preference="direct_solve", is_sparse_problem=False, is_hermitian=True)
    assert step_population(cands, H, None, strat, know, eng) == C
    w, E = sla.eigh(H)                                              # AMS:161
    for c, v0 in zip(cands, v_before):
        best = int(np.argmax(np.abs(v0.conj().T @ E)))             # AMS:165-169
        assert abs(c.lambda_k - w[best]) <= 1e-12 * np.abs(w).max()
        assert isinstance(c.lambda_k, (float, np.floating))         # a real eigenvalue, like the reference's (AMS:171)
        ph = np.vdot(c.v_k, E[:, best]); ph /= abs(ph)
        assert np.abs(c.v_k * ph - E[:, best]).max() <= 1e-10
        assert c.state == MockCandidate.State.CONVERGED and c.stuck_counter == 0 and c.w_k == 1.0
        r = np.linalg.norm(H @ c.v_k - c.lambda_k * c.v_k)
        assert abs(c.residual_k - r) <= 1e-12 and c.residual_k <= 1e-12 * np.linalg.norm(H)
        assert len(c.residual_history) == 2

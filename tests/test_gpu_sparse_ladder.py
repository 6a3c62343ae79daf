"""The retry ladder on SPARSE eigenproblems (BASELINE config 5, eigen half; AMS:43-104, 57, 99-102).

With the Rayleigh-quotient shift of a random start vector inside the spectrum, GMRES(20) x 50 stagnates -- in scipy exactly as on
the device -- and the reference switches to its direct solver (SuperLU) at attempt 0.  The device mirrors that with the batched
dense LU while the order allows it (n <= 8192: the dense form is attached to the resident CSR copy on first need); beyond that
order there is no device direct solver, the try counts as failed and the candidate is re-initialised.  Both outcomes are pinned
here against the oracle, side by side."""
import random
import warnings

import numpy as np
import pytest

from mock_candidate import MockCandidate, ProblemType
from oracle import maus_oracle as mo
from parity import anorm, assert_scalar_close, vec_err_up_to_phase

pytestmark = pytest.mark.gpu

STRAT = dict(overall_psi_aggression_factor=1.0, max_psi_retries=3, current_convergence_threshold=1e-10)
KNOW = dict(local_solver_preference="iterative_gmres", is_sparse_problem=True, is_hermitian=False)


def _population(n, C, seed):
    from adaptive_matrix_solver_b200.workloads import k5_sparse
    A = k5_sparse(n, seed=seed)
    np.random.seed(seed); random.seed(seed)
    cands = [MockCandidate(A, ProblemType.EIGENVALUE, n) for _ in range(C)]
    for c in cands:
        c.alpha_local_step = 0.5
    return A, cands


@pytest.mark.timeout(600)
def test_sparse_eigen_gmres_stagnates_then_direct_fallback_matches_the_reference():
    import adaptive_matrix_solver_b200 as pkg
    from adaptive_matrix_solver_b200 import step_population
    n, C = 1200, 3
    A, cands = _population(n, C, 4)
    oracles = [c.to_oracle() for c in cands]
    traces = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for o in oracles:
            tr = []
            mo.candidate_step(o, A, None, STRAT, KNOW, gmres_mode="shim", trace=tr)
            traces.append(tr)
    assert all(tr and tr[0][1] != 0 for tr in traces)              # scipy's GMRES did not converge either (info != 0)
    eng = pkg.MausEngine(0)
    step_population(cands, A, None, STRAT, KNOW, eng)
    assert eng.is_sparse and eng.has_dense_form                    # CSR kept for the matvecs, dense form attached for the LU
    floor = 4e-13 * anorm(A)
    for c, o in zip(cands, oracles):
        assert c.local_psi_retries_needed == o.local_psi_retries_needed == 0       # fallback succeeded at attempt 0 (AMS:99-102)
        assert c.stuck_counter == o.stuck_counter and c.state.value == o.state
        assert_scalar_close(c.lambda_k, o.lambda_k, floor, "lambda")
        assert_scalar_close(c.residual_k, o.residual_k, 1e-9 * anorm(A), "residual")
        assert vec_err_up_to_phase(c.v_k, o.v_k) <= 1e-8
    # the next generation still multiplies with the sparse copy and can fall back again
    step_population(cands, A, None, STRAT, KNOW, eng)
    assert all(np.isfinite(c.residual_k) for c in cands)
    eng.close()


@pytest.mark.timeout(600)
def test_beyond_the_lu_limit_the_ladder_fails_where_the_reference_falls_back(monkeypatch):
    """Documented deviation: sparse order above the batched-LU limit (emulated here by lowering the limit below
    n = 1200).  Reference: SuperLU fallback succeeds, the candidate takes the success branch.  Device: every attempt fails,
    the candidate takes the RuntimeError branch of AMS:287-293 (stuck + 1, weight x 0.001, alpha halved, re-initialised)."""
    import adaptive_matrix_solver_b200 as pkg
    from adaptive_matrix_solver_b200 import population, step_population
    monkeypatch.setattr(population, "LU_MAX_N", 1000)
    n, C = 1200, 2
    A, cands = _population(n, C, 6)
    oracles = [c.to_oracle() for c in cands]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for o in oracles:
            mo.candidate_step(o, A, None, STRAT, KNOW, gmres_mode="shim")
    eng = pkg.MausEngine(0)
    step_population(cands, A, None, STRAT, KNOW, eng)
    eng.close()
    for c, o in zip(cands, oracles):
        assert o.stuck_counter == 0 and o.w_k == 0.01 and o.num_resets == 0          # reference: success branch
        assert c.stuck_counter == 1 and c.w_k == 0.01 * 0.001                          # device: failure branch
        assert len(c.residual_history) == o.history_len + 1                           # + the re-initialisation entry (AMS:142-143)
        assert np.isfinite(c.residual_k) and abs(np.linalg.norm(c.v_k) - 1.0) < 1e-12

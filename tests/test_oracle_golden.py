"""Pins oracle/maus_oracle.py to the REAL reference: every golden step (recorded by oracle/gen_golden.py from
/root/reference/Adaptive_Matrix_Solver_0.1.py under fixed seeds) is replayed through the oracle and must match
bit for bit (same numpy/scipy ops in the same order; tolerances only where BLAS threading may reorder sums)."""
import random
import warnings

import numpy as np
import pytest
from threadpoolctl import threadpool_limits

from golden_io import Golden, NAMES
from oracle import maus_oracle as mo


def _make_state(g, i):
    s = g.side("before", i)
    c = mo.CandState(problem_type=g.problem_type, N=g.n)
    c.lambda_k = s["lam"]; c.v_k = None if s["v"] is None else s["v"].copy()
    c.x_k = None if s["x"] is None else s["x"].copy()
    c.state = s["state"]; c.w_k = s["w"]; c.residual_k = s["res"]; c.prev_residual = s["prev"]
    c.alpha_local_step = s["alpha"]; c.stuck_counter = s["stuck"]; c.local_psi_retries_needed = s["retries"]
    c.num_resets = s["resets"]; c.history_len = s["hist"]
    if g.problem_type == 3:
        c.M_rows, c.M_cols = g.A.shape
        c.u_k = s["u"].copy(); c.right_v_k = s["rv"].copy(); c.sigma_k = s["sigma"]
        c.v_k = None; c.x_k = None
    return c


def _close(a, b, rtol):
    a = np.asarray(a); b = np.asarray(b)
    both_nan = np.isnan(a) & np.isnan(b)
    return bool(np.all(both_nan | (a == b) | (np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)))))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_replays_reference_steps(name):
    g = Golden(name)
    assert g.n_steps > 0
    # LAPACK / BLAS are deterministic for a fixed thread count; allow a few ulp for thread-count differences
    rtol = 1e-12
    worst = 0.0
    # goldens were generated with one BLAS thread (OPENBLAS_NUM_THREADS=1); near-singular inverse-iteration
    # solves amplify summation-order differences, so the replay pins the pool to one thread as well
    with warnings.catch_warnings(), threadpool_limits(limits=1):
        warnings.simplefilter("ignore")
        for i in range(g.n_steps):
            c = _make_state(g, i)
            seed = int(g.z["seed"][i])
            np.random.seed(seed % (2 ** 32)); random.seed(seed)
            mo.candidate_step(c, g.A, g.b, g.strat(i), g.know(i), problem_matrix_ctor=g.ctor_matrix(i),
                              gmres_mode=g.gmres_mode)
            a = g.side("after", i)
            assert c.state == a["state"], (name, i)
            assert c.stuck_counter == a["stuck"], (name, i)
            assert c.local_psi_retries_needed == a["retries"], (name, i)
            assert c.num_resets == a["resets"], (name, i)
            assert c.history_len == a["hist"], (name, i)
            assert c.w_k == a["w"], (name, i)
            assert complex(c.alpha_local_step) == complex(a["alpha"]), (name, i)
            assert isinstance(c.alpha_local_step, (complex, np.complexfloating)) == isinstance(
                a["alpha"], (complex, np.complexfloating)), (name, i)
            assert _close(c.residual_k, a["res"], 1e-9), (name, i, c.residual_k, a["res"])
            assert _close(c.prev_residual, a["prev"], 0.0), (name, i)
            if g.problem_type == 1:
                assert _close(c.lambda_k, a["lam"], rtol), (name, i, c.lambda_k, a["lam"])
                assert _close(c.v_k, a["v"], 1e-9), (name, i)
            elif g.problem_type == 2:
                assert _close(c.x_k, a["x"], 1e-9), (name, i)
            else:
                assert _close(float(np.real(c.sigma_k)), a["sigma"], rtol), (name, i)
                assert _close(c.u_k, a["u"], 1e-9) and _close(c.right_v_k, a["rv"], 1e-9), (name, i)


def test_psi_magnitude_matches_reference_formula():
    # AMS:44 with the constants of AMS:16, 224
    base = mo.PSI_EPSILON_BASE * 50.0
    assert mo.psi_magnitude(base, 3, 2) == base * (10 ** 1.5) * (10 ** (2 / 3.0))
    assert isinstance(mo.psi_magnitude(base, 0, 0), np.complexfloating)


def test_ladder_raises_runtime_error_when_all_attempts_fail():
    A = np.full((4, 4), np.nan, dtype=np.complex128)
    with pytest.raises(RuntimeError):
        mo.inverse_iterate_solve(A, np.ones(4, dtype=np.complex128), 0, N=4, base_psi_epsilon=mo.PSI_EPSILON_BASE,
                                 max_attempts=3)


def test_gmres_as_shipped_falls_back_to_direct():
    rng = np.random.default_rng(0)
    A = rng.random((6, 6)) + 1j * rng.random((6, 6)) + 6 * np.eye(6)
    b = np.ones(6, dtype=np.complex128)
    tr = []
    x, att = mo.inverse_iterate_solve(A, b, 0, N=6, base_psi_epsilon=mo.PSI_EPSILON_BASE, max_attempts=5,
                                      preferred_method="iterative_gmres", gmres_mode="as_shipped", trace=tr)
    assert att == 0 and tr == []            # gmres never ran (TypeError path), direct solve answered
    assert np.allclose(A @ x, b)
    x2, _ = mo.inverse_iterate_solve(A, b, 0, N=6, base_psi_epsilon=mo.PSI_EPSILON_BASE, max_attempts=5,
                                     preferred_method="iterative_gmres", gmres_mode="shim", trace=tr)
    assert tr and tr[0][0] == "gmres" and tr[0][1] == 0

"""Set-up diagnostics on the device: drop-in for ``MAUS_Solver._diagnose_matrix_initial`` (AMS:374-404).

The reference diagnoses the problem matrix once, in ``MAUS_Solver.__init__``: two ``np.allclose`` passes (Hermitian /
complex-symmetric), ``np.count_nonzero`` and ``np.linalg.cond`` -- a full SVD that dominates start-up from n ~ 2000 on (minutes at
n = 8192) -- and then picks one of three strategies from ``cond > 1e12`` / ``cond > 1e6`` (``_set_initial_strategy``,
AMS:405-421).  ``diagnose_matrix_initial`` returns the SAME dictionary, computed by libmaus_b200 on the matrix it uploads for
the candidate steps anyway:

  * sparsity, ``is_hermitian``, ``is_complex_symmetric``: exact (same element-wise ``isclose`` formula, one device pass);
  * ``condition_number``: sigma_max / sigma_min from power / inverse iteration (inverse iteration runs on the batched LU of the
    hot path).  It is a LOWER bound of cond_2, tight to a few per cent; the three-way decision is the reference's unless the
    true condition number sits within that margin of 1e6 / 1e12 / 1e15 (documented deviation; ``margin`` below is reported so a
    caller can fall back to the exact SVD in that case).

Sparse inputs follow the reference's branch unchanged (AMS:386-395: densified property tests for n^2 <= 1e7, no condition
number -> ``inf``).  ``install_diagnostics(ams, engine)`` rebinds the method the way ``install_dropin`` rebinds the solver.
"""
import numpy as np

from .constants import MAX_PSI_ATTEMPTS

COND_THRESHOLDS = (1e6, 1e12, 1e15)          # Fragile / Critical (AMS:406, 410) and the singularity flag (AMS:401)


def _is_sparse(A):
    try:
        import scipy.sparse as sp
        return sp.issparse(A)
    except Exception:  # pragma: no cover
        return False


def diagnose_matrix_initial(engine, matrix, power_iters=40, inverse_iters=8, keep_resident=True):
    """AMS:374-404 on the device.  Returns the reference's ``diag_info`` dict plus ``cond_margin`` (distance of the estimate to
    the nearest decision threshold, as a factor >= 1)."""
    diag_info = {'is_hermitian': False, 'is_complex_symmetric': False, 'is_sparse_init': False,
                 'condition_number': np.inf, 'is_singular': False}
    if _is_sparse(matrix):                                                               # AMS:386-395, unchanged (host)
        diag_info['is_sparse_init'] = True
        try:
            if matrix.shape[0] == matrix.shape[1] and matrix.shape[0] * matrix.shape[1] <= 1e7:
                D = np.asarray(matrix.todense())
                if np.allclose(D, D.conj().T):
                    diag_info['is_hermitian'] = True
                if np.allclose(D, D.T):
                    diag_info['is_complex_symmetric'] = True
        except Exception:
            pass
        return diag_info
    if not isinstance(matrix, np.ndarray):
        return diag_info
    square = matrix.ndim == 2 and matrix.shape[0] == matrix.shape[1] and matrix.size > 0
    if not square:
        diag_info['is_sparse_init'] = (np.count_nonzero(matrix) / matrix.size) < 0.25 if matrix.size > 0 else False
        return diag_info                                                                 # rectangular (SVD problems): no cond, no symmetry
    engine.set_matrix(matrix)
    nz, herm, sym = engine.diag_dense()
    diag_info['is_sparse_init'] = (nz / matrix.size) < 0.25                               # AMS:381
    diag_info['is_hermitian'] = herm                                                     # AMS:384
    diag_info['is_complex_symmetric'] = sym                                              # AMS:385
    if not diag_info['is_sparse_init']:                                                  # AMS:397-402
        smax, smin, st = engine.cond2_estimate(power_iters, inverse_iters)
        cond = np.inf if (st != 0 or not smin > 0.0 or not np.isfinite(smax)) else smax / smin
        diag_info['condition_number'] = cond
        diag_info['is_singular'] = bool(np.isinf(cond) or cond > 1e15)
        diag_info['cond_margin'] = (np.inf if not np.isfinite(cond) else
                                    min(max(cond / t, t / cond) for t in COND_THRESHOLDS))
    return diag_info


def initial_strategy(diag_info, problem_type_name, convergence_tolerance=1e-8):
    """``_set_initial_strategy`` (AMS:405-421) as a pure function of the diagnosis: returns (strat_params updates,
    problem_knowledge updates).  ``problem_type_name``: 'EIGENVALUE' | 'SOLVE_LINEAR_SYSTEM' | 'SVD'."""
    cond = diag_info['condition_number']
    strat = {'overall_psi_aggression_factor': 1.0, 'max_psi_retries': MAX_PSI_ATTEMPTS,
             'current_convergence_threshold': convergence_tolerance}
    know = {'true_matrix_is_singular': bool(diag_info.get('is_singular', False))}
    if cond > 1e12:
        know['numerical_stability_state'] = 'Critical'; strat['overall_psi_aggression_factor'] = 50.0
        strat['max_psi_retries'] = MAX_PSI_ATTEMPTS * 2; strat['current_convergence_threshold'] = 1e-2
        know['local_solver_preference'] = 'iterative_gmres'
    elif cond > 1e6:
        know['numerical_stability_state'] = 'Fragile'; strat['overall_psi_aggression_factor'] = 10.0
        know['local_solver_preference'] = 'iterative_gmres'; strat['current_convergence_threshold'] = 1e-4
    else:
        know['numerical_stability_state'] = 'Stable'; know['local_solver_preference'] = 'direct_solve'
    if problem_type_name == 'SOLVE_LINEAR_SYSTEM' and diag_info.get('is_singular', False):
        know['true_matrix_is_singular'] = True; know['local_solver_preference'] = 'iterative_gmres'
        strat['overall_psi_aggression_factor'] = max(strat['overall_psi_aggression_factor'], 20.0)
    if problem_type_name == 'SVD':
        if know['numerical_stability_state'] == 'Stable':
            strat['overall_psi_aggression_factor'] = max(strat['overall_psi_aggression_factor'], 2.0)
        strat['current_convergence_threshold'] = max(1e-5, convergence_tolerance)
    return strat, know


def install_diagnostics(ams_module, engine, **kw):
    """Rebind ``MAUS_Solver._diagnose_matrix_initial`` (looked up on the class at AMS:345) to the device version."""
    def _diagnose(self, matrix):
        return diagnose_matrix_initial(engine, matrix, **kw)
    ams_module.MAUS_Solver._diagnose_matrix_initial = _diagnose
    return ams_module

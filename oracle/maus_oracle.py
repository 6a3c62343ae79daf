"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MAUS candidate-step hot path.

This file is a numpy/scipy *restatement* of the reference's algorithm for ONE path: the per-candidate
Psi-regularised shifted inverse-iteration step (``InverseIterateSolver.solve`` + the eigen / linear-system
branch of ``SolutionCandidate.update_solution_step``, plus the SVD power sweep
of the same method -- the first 'next' row of SURVEY.md section 8f).  AMS = /root/reference/Adaptive_Matrix_Solver_0.1.py.
It is the checker for the CUDA path; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package never does.

Pinning: the reference ships NO tests / golden vectors for this path, so the oracle is
pinned against traces produced by running the real reference in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; checked bit-for-bit by tests/test_oracle_golden.py).

Third-party arithmetic, exactly as the reference uses it (AMS:57,59,89; un-vendored, no pin in the reference;
the build container has numpy 2.3.5 / scipy 1.18.1 / OpenBLAS 0.3.30): ``scipy.linalg.solve`` (LAPACK zgesv),
``scipy.sparse.linalg.spsolve`` (SuperLU) and ``scipy.sparse.linalg.gmres``.  ``gmres_mode`` selects how AMS:89
behaves: ``"shim"`` (tol forwarded as rtol -- GMRES really runs) or ``"as_shipped"`` (TypeError -> caught at
AMS:98 -> direct fallback), see SURVEY.md section 0.3.

The operation ORDER below deliberately follows the reference line by line so that, given the same global
``np.random`` / ``random`` state, results are bit-identical to the reference's.
"""
from dataclasses import dataclass, field
import random as _pyrandom

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --- constants, AMS:16-26 -------------------------------------------------------------------------------
PSI_EPSILON_BASE = np.complex128(1e-20)      # AMS:16
ALPHA_V_INITIAL = np.complex128(0.01)        # AMS:17
MAX_PSI_ATTEMPTS = 25                        # AMS:18
MAX_STUCK_FOR_RETIREMENT = 8                 # AMS:19
SIGMA_SIMILARITY_TOL_ABS = 1e-6              # AMS:23
CONVERGENCE_RESIDUAL_TOL = 1e-8              # AMS:25

# problem types (AMS:10-13) and candidate states (AMS:109-110) as plain ints
EIGENVALUE, SOLVE_LINEAR_SYSTEM, SVD = 1, 2, 3
EXPLORING, REFINING, STUCK, CONVERGED, RETIRED = 1, 2, 3, 4, 5


# --- InverseIterateSolver, AMS:30-104 -------------------------------------------------------------------
def psi_magnitude(base_psi_epsilon, num_psi_attempts, candidate_stuck_counter):
    """AMS:44."""
    return base_psi_epsilon * (10 ** (num_psi_attempts / 2.0)) * (10 ** (candidate_stuck_counter / 3.0))


def regulariser(N, psi, is_sparse, dtype, rand=None):
    """AMS:46-50.  ``rand(N, N)`` defaults to the global ``np.random.rand`` exactly like the reference."""
    if is_sparse:
        return sp.identity(N, dtype=dtype, format="csc") * psi
    rand = np.random.rand if rand is None else rand
    random_perturb = (rand(N, N) - 0.5 + 1j * (rand(N, N) - 0.5)) * psi * 0.15
    return psi * np.eye(N, dtype=dtype) + random_perturb


def jacobi_preconditioner(H_solve, N, candidate_stuck_counter, is_sparse):
    """AMS:64-86.  Returns the inverse diagonal as the reference builds it, or None."""
    if not (candidate_stuck_counter > 1 and N > 0):
        return None
    diag_H = H_solve.diagonal()
    if diag_H.size != N:
        return None
    with np.errstate(divide="ignore", invalid="ignore"):
        inv_diag_H = 1.0 / diag_H
    if np.all(np.isfinite(inv_diag_H)) and np.all(np.abs(diag_H) > 1e-12):
        return sp.diags(inv_diag_H, format="csc") if is_sparse else np.diag(inv_diag_H)
    return None


def inverse_iterate_solve(A_target, b_rhs, candidate_stuck_counter, *, N, base_psi_epsilon, max_attempts,
                          preferred_method="direct_solve", is_sparse=False, gmres_mode="shim", rand=None,
                          trace=None):
    """AMS:39-104.  Returns (result_vec, num_psi_attempts) or raises RuntimeError (AMS:104)."""
    fallback_method = "iterative_gmres" if preferred_method == "direct_solve" else "direct_solve"  # AMS:36
    num_psi_attempts = 0
    method = preferred_method
    while num_psi_attempts < max_attempts:                                                     # AMS:43
        psi = psi_magnitude(base_psi_epsilon, num_psi_attempts, candidate_stuck_counter)
        H_solve = A_target + regulariser(N, psi, is_sparse, A_target.dtype, rand)               # AMS:52
        try:
            if method == "direct_solve":
                if is_sparse:
                    result_vec = spla.spsolve(H_solve.tocsc(), b_rhs)                          # AMS:57
                else:
                    result_vec = sla.solve(H_solve, b_rhs, assume_a="general")                 # AMS:59
            elif method == "iterative_gmres":
                x0_init = b_rhs if b_rhs.shape == H_solve.shape[1:] else np.zeros_like(b_rhs)   # AMS:61
                M = jacobi_preconditioner(H_solve, N, candidate_stuck_counter, is_sparse)
                if gmres_mode == "as_shipped":
                    raise TypeError("gmres() got an unexpected keyword argument 'tol'")        # AMS:89 on scipy>=1.14
                result_vec, info = spla.gmres(H_solve, b_rhs, x0=x0_init, rtol=1e-8, maxiter=50, M=M)
                if trace is not None:
                    trace.append(("gmres", info, M is not None))
                if info != 0:
                    raise np.linalg.LinAlgError(f"GMRES did not converge cleanly (info={info}).")  # AMS:90
            else:
                raise ValueError(f"Unknown solver method: {method}")
            if not np.all(np.isfinite(result_vec)):
                raise ValueError("Solution vector not finite after solve.")                     # AMS:94-95
            return result_vec, num_psi_attempts                                                 # AMS:97
        except (np.linalg.LinAlgError, ValueError, TypeError):
            if method == preferred_method and preferred_method != fallback_method and num_psi_attempts == 0:
                method = fallback_method                                                        # AMS:99-102
                num_psi_attempts = 0
                continue
            num_psi_attempts += 1                                                               # AMS:103
    raise RuntimeError(f"InverseIterateSolver failed all {max_attempts} attempts.")            # AMS:104


# --- candidate state, AMS:113-126 -----------------------------------------------------------------------
@dataclass
class CandState:
    problem_type: int
    N: int
    lambda_k: object = None
    v_k: object = None
    x_k: object = None
    state: int = EXPLORING
    w_k: float = 0.01
    residual_k: float = float("inf")
    prev_residual: float = float("inf")
    u_k: object = None
    right_v_k: object = None
    sigma_k: object = None
    M_rows: int = 0
    M_cols: int = 0
    alpha_local_step: object = ALPHA_V_INITIAL
    stuck_counter: int = 0
    local_psi_retries_needed: int = 0
    num_resets: int = 0
    history_len: int = 0
    extra: dict = field(default_factory=dict)

    def copy(self):
        c = CandState(**{k: getattr(self, k) for k in self.__dataclass_fields__ if k != "extra"})
        if c.v_k is not None:
            c.v_k = np.array(c.v_k, copy=True)
        if c.x_k is not None:
            c.x_k = np.array(c.x_k, copy=True)
        if c.u_k is not None:
            c.u_k = np.array(c.u_k, copy=True)
        if c.right_v_k is not None:
            c.right_v_k = np.array(c.right_v_k, copy=True)
        return c


def _rand_vec_init(N):
    return (np.random.rand(N) + 1j * np.random.rand(N)).astype(np.complex128)   # AMS:130


def _norm_rand_vec(v):
    # AMS:131 -- note the norm is evaluated twice, and the fallback draws two fresh vectors
    if np.linalg.norm(v) > 1e-10:
        return v / np.linalg.norm(v)
    return _rand_vec_init(v.shape[0]) / np.linalg.norm(_rand_vec_init(v.shape[0]))


def initialize_random_solution(c):
    """AMS:129-143 (eigen / linear branches)."""
    if c.problem_type == EIGENVALUE:
        c.v_k = _norm_rand_vec(_rand_vec_init(c.N))
        c.lambda_k = (_pyrandom.random() * 5 - 2.5 + 1j * (_pyrandom.random() * 5 - 2.5))
    elif c.problem_type == SOLVE_LINEAR_SYSTEM:
        c.x_k = _norm_rand_vec(_rand_vec_init(c.N)) * _pyrandom.uniform(0.1, 10.0)
    elif c.problem_type == SVD:
        c.u_k = _norm_rand_vec(_rand_vec_init(c.M_rows))
        c.right_v_k = _norm_rand_vec(_rand_vec_init(c.M_cols))
        c.sigma_k = 1.0
    c.history_len += 1


def adapt_alpha_and_state(c):
    """AMS:306-316."""
    if c.prev_residual > 1e-10:
        if c.residual_k < c.prev_residual * 0.9:
            c.alpha_local_step = min(c.alpha_local_step * 1.1, 1.0)
            if c.state != CONVERGED:
                c.state = REFINING
        elif c.residual_k > c.prev_residual * 1.5 and c.prev_residual > 1e-5:
            c.alpha_local_step = max(c.alpha_local_step * 0.5, 1e-6)
            if c.state != CONVERGED:
                c.state = STUCK
        else:
            c.alpha_local_step = max(c.alpha_local_step * 0.95, 1e-6)
            if c.state not in (CONVERGED, STUCK, RETIRED):
                c.state = EXPLORING


def convergence_test(c, current_conv_tol):
    """AMS:318-331."""
    if c.problem_type == EIGENVALUE:
        params = (c.lambda_k, c.v_k)
    elif c.problem_type == SOLVE_LINEAR_SYSTEM:
        params = (c.x_k,)
    else:
        params = (c.sigma_k, c.u_k, c.right_v_k)
    finite = True
    for p in params:
        if p is None:
            finite = False
            break
        if isinstance(p, np.ndarray):
            if not np.all(np.isfinite(p)):
                finite = False
                break
        elif not np.isfinite(p):
            finite = False
            break
    if c.residual_k < current_conv_tol and finite:
        c.state = CONVERGED
        c.w_k = 1.0
        c.stuck_counter = 0
        c.alpha_local_step = 0.0


def candidate_step(c, current_matrix_A, b_vector, strat_params, global_knowledge, problem_matrix_ctor=None,
                   gmres_mode="shim", rand=None, trace=None):
    """The eigen (non-Hermitian) / linear-system branch of update_solution_step, AMS:145-153, 224-225,
    256-299, 303-331.  Mutates and returns ``c``.  ``problem_matrix_ctor`` is the ctor-time matrix the
    reference uses for the residual (AMS:118, 295); defaults to ``current_matrix_A``."""
    A_res_calc = current_matrix_A if problem_matrix_ctor is None else problem_matrix_ctor
    c.prev_residual = c.residual_k                                                             # AMS:147
    aggr = strat_params.get("overall_psi_aggression_factor", 1.0)
    max_retries = strat_params.get("max_psi_retries", MAX_PSI_ATTEMPTS)
    pref = global_knowledge.get("local_solver_preference", "direct_solve")
    is_sparse = global_knowledge.get("is_sparse_problem", False)
    N = c.N
    if c.problem_type == SVD:
        return _svd_step(c, current_matrix_A, A_res_calc, strat_params)
    solver_kw = dict(N=N, base_psi_epsilon=PSI_EPSILON_BASE * aggr, max_attempts=max_retries,
                     preferred_method=pref, is_sparse=is_sparse, gmres_mode=gmres_mode, rand=rand, trace=trace)

    if c.problem_type == EIGENVALUE:
        if np.linalg.norm(c.v_k) < 1e-10:                                                      # AMS:259-263
            c.v_k = (np.random.rand(N) + 1j * np.random.rand(N))
            c.v_k /= np.linalg.norm(c.v_k)
            c.stuck_counter += 1
            c.num_resets += 1
        denom = np.vdot(c.v_k, c.v_k)                                                          # AMS:264
        if np.abs(denom) < 1e-12:
            c.lambda_k = complex(0.0, 0.0)
        else:
            c.lambda_k = np.vdot(c.v_k, current_matrix_A @ c.v_k) / denom                      # AMS:268
        eye = sp.eye(N, dtype=current_matrix_A.dtype) if is_sparse else np.eye(N, dtype=current_matrix_A.dtype)
        target = current_matrix_A - c.lambda_k * eye                                           # AMS:270
        rhs = c.v_k
        main_ref = c.v_k
    elif c.problem_type == SOLVE_LINEAR_SYSTEM:
        target = current_matrix_A
        rhs = b_vector
        main_ref = c.x_k
    else:
        raise ValueError("oracle covers the EIGENVALUE / SOLVE_LINEAR_SYSTEM branches only")
    try:
        new_vec_raw, c.local_psi_retries_needed = inverse_iterate_solve(target, rhs, c.stuck_counter, **solver_kw)
        if c.problem_type == EIGENVALUE:
            c.v_k = (1.0 - c.alpha_local_step) * c.v_k + c.alpha_local_step * new_vec_raw       # AMS:280
            norm_v_k = np.linalg.norm(c.v_k)
            if norm_v_k > 1e-10:
                c.v_k /= norm_v_k
            else:
                c.v_k = (np.random.rand(N) + 1j * np.random.rand(N)) / np.sqrt(N)
        else:
            c.x_k = (1.0 - c.alpha_local_step) * main_ref + c.alpha_local_step * new_vec_raw    # AMS:285
        c.stuck_counter = max(0, c.stuck_counter - 1)                                          # AMS:286
    except (RuntimeError, ValueError):
        c.stuck_counter += 1
        c.w_k *= 0.001
        c.alpha_local_step = max(c.alpha_local_step * 0.5, 1e-6)                               # AMS:288-289
        if c.stuck_counter >= MAX_STUCK_FOR_RETIREMENT:
            c.state = RETIRED
            c.num_resets += 1
        else:
            c.state = STUCK
            initialize_random_solution(c)

    if c.problem_type == EIGENVALUE:                                                           # AMS:295-299
        c.residual_k = np.linalg.norm(A_res_calc @ c.v_k - c.lambda_k * c.v_k)
    else:
        c.residual_k = np.linalg.norm(A_res_calc @ c.x_k - b_vector)
    c.history_len += 1                                                                         # AMS:303-304
    adapt_alpha_and_state(c)
    convergence_test(c, strat_params.get("current_convergence_threshold", CONVERGENCE_RESIDUAL_TOL))
    return c


def _svd_step(c, A, A_res_calc, strat_params):
    """SVD branch of update_solution_step: one power sweep u = Av/||.||, v = A^H u/||.|| (AMS:227-255), residual
    (AMS:300-301), then the common tail (AMS:303-331).  Never calls the inverse-iteration solver."""
    Mr, Mc = c.M_rows, c.M_cols
    try:
        if np.linalg.norm(c.right_v_k) < 1e-10:                                                 # AMS:229-232
            c.right_v_k = (np.random.rand(Mc) + 1j * np.random.rand(Mc))
            c.right_v_k /= np.linalg.norm(c.right_v_k)
            c.stuck_counter += 1
            c.num_resets += 1
            raise ValueError("SVD right_v_k collapsed.")
        temp_u_k = A @ c.right_v_k                                                              # AMS:233
        c.sigma_k = np.linalg.norm(temp_u_k)
        c.u_k = temp_u_k / (c.sigma_k if c.sigma_k > 1e-10 else 1.0)
        if np.linalg.norm(c.u_k) < 1e-10:                                                       # AMS:236-239
            c.u_k = (np.random.rand(Mr) + 1j * np.random.rand(Mr))
            c.u_k /= np.linalg.norm(c.u_k)
            c.stuck_counter += 1
            c.num_resets += 1
            raise ValueError("SVD u_k collapsed.")
        temp_v_k = A.conj().T @ c.u_k                                                           # AMS:240
        c.sigma_k = max(c.sigma_k, np.linalg.norm(temp_v_k))
        c.right_v_k = temp_v_k / (np.linalg.norm(temp_v_k) if np.linalg.norm(temp_v_k) > 1e-10 else 1.0)
        if c.sigma_k < SIGMA_SIMILARITY_TOL_ABS / 100:                                          # AMS:243-247
            c.residual_k = strat_params.get("current_convergence_threshold", 1e-6) * 0.1
            c.state = CONVERGED
            c.stuck_counter = 0
            if np.linalg.norm(c.u_k) < 1e-10:
                c.u_k = np.ones(Mr, dtype=np.complex128) / np.sqrt(Mr)
            if np.linalg.norm(c.right_v_k) < 1e-10:
                c.right_v_k = np.ones(Mc, dtype=np.complex128) / np.sqrt(Mc)
        else:
            c.stuck_counter = max(0, c.stuck_counter - 1)                                       # AMS:248
    except (RuntimeError, ValueError, np.linalg.LinAlgError):                                   # AMS:249-255
        c.stuck_counter += 1
        c.w_k *= 0.001
        c.alpha_local_step *= 0.5
        c.state = STUCK
        if c.stuck_counter >= MAX_STUCK_FOR_RETIREMENT:
            c.state = RETIRED
        c.u_k = (np.random.rand(Mr) + 1j * np.random.rand(Mr)) / np.sqrt(Mr)
        c.right_v_k = (np.random.rand(Mc) + 1j * np.random.rand(Mc)) / np.sqrt(Mc)
        c.sigma_k = 1.0
    c.residual_k = (np.linalg.norm(A_res_calc @ c.right_v_k - c.sigma_k * c.u_k)
                    + np.linalg.norm(A_res_calc.conj().T @ c.u_k - c.sigma_k * c.right_v_k))    # AMS:301
    c.history_len += 1
    adapt_alpha_and_state(c)
    convergence_test(c, strat_params.get("current_convergence_threshold", CONVERGENCE_RESIDUAL_TOL))
    return c


# --- deterministic single-step pieces used by the GPU parity tests --------------------------------------
def rayleigh_quotient(A, v):
    """AMS:264-268."""
    denom = np.vdot(v, v)
    if np.abs(denom) < 1e-12:
        return complex(0.0, 0.0)
    return np.vdot(v, A @ v) / denom


def shifted_solve_dense(A, lam, psi, rhs, R=None):
    """x = (A - lam I + psi I + R)^-1 rhs with the reference's LAPACK call (AMS:50-52, 59, 270)."""
    N = A.shape[0]
    T = A - lam * np.eye(N, dtype=A.dtype)
    H = T + (psi * np.eye(N, dtype=A.dtype) + (0 if R is None else R))
    return sla.solve(H, rhs, assume_a="general")


def mix_normalise(v, x, alpha):
    """AMS:280-282 (eigen)."""
    v2 = (1.0 - alpha) * v + alpha * x
    nv = np.linalg.norm(v2)
    return (v2 / nv if nv > 1e-10 else v2), nv


def residual_eigen(A_ctor, v, lam):
    """AMS:297."""
    return np.linalg.norm(A_ctor @ v - lam * v)


def residual_linear(A_ctor, x, b):
    """AMS:299."""
    return np.linalg.norm(A_ctor @ x - b)

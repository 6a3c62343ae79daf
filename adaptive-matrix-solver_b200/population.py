"""Seam B: one batched GPU generation behind the reference's candidate objects.

``step_population(candidates, M, b, strat_params, problem_knowledge, engine)`` is the drop-in for the loop at
AMS:574-576::

    for candidate in self.candidates:
        if candidate.state not in [CONVERGED, RETIRED]:
            candidate.update_solution_step(self.M, self.b, self.strat_params, self.problem_knowledge)

It performs, for every live EIGENVALUE (non-Hermitian) / SOLVE_LINEAR_SYSTEM candidate, exactly what
``SolutionCandidate.update_solution_step`` does (AMS:145-153, 224-225, 256-299, 303-331) and writes back every field
the reference mutates, so ``_update_global_diagnostics`` / ``_adjust_global_strategy`` / ``_manage_candidates`` run
unmodified afterwards.  All numerics (Rayleigh quotient, shifted solve, mix, normalise, residual) run in ONE fused
CUDA call for the whole population; only the rare failures walk the Psi ladder (AMS:43-104) through the granular
calls.  SVD candidates run the batched power sweep of ``svd.cu``; dense Hermitian-eigen candidates
share one device eigendecomposition (8f-3); sparse Hermitian ones (eigsh) are passed to the reference method untouched.  With
``engine.enable_row_sharding`` a sparse problem on the GMRES path runs on the row-sharded operator (BASELINE config 5).

Deliberate deviations (DESIGN.md section "Deviations"): the dense Psi perturbation (AMS:49) comes from a
counter-based device RNG, not from the global numpy stream; host RNG draws therefore happen only for the
re-initialisations (AMS:260, 283, 293), in candidate order.
"""
import numpy as np

from . import _abi
from .constants import (PSI_EPSILON_BASE, MAX_PSI_ATTEMPTS, MAX_STUCK_FOR_RETIREMENT, CONVERGENCE_RESIDUAL_TOL,
                        SIGMA_SIMILARITY_TOL_ABS, LU_MAX_N, psi_magnitude)
from .solver import GpuInverseIterateSolver

_METHOD = {'direct_solve': _abi.METHOD_LU, 'iterative_gmres': _abi.METHOD_GMRES}


def _is_sparse(A):
    try:
        import scipy.sparse as sp
        return sp.issparse(A)
    except Exception:  # pragma: no cover
        return False


class _MatrixCache:
    """Uploads a matrix only when the object behind the reference's ``self.M`` changes (AMS:646 swaps it)."""

    def __init__(self):
        self.obj = [None, None]
        self.form = [None, None]
        self.epoch = [None, None]

    def ensure(self, engine, A, slot, form):
        # the engine counts every upload per slot (``matrix_epoch``): a Seam-A solve or any other direct ``set_matrix`` on the
        # shared engine between two generations (AMS:224 pass-through candidates) invalidates what this cache believes is resident
        epoch = getattr(engine, "matrix_epoch", None)
        if self.obj[slot] is A and self.form[slot] == form and (epoch is None or self.epoch[slot] == epoch[slot]):
            return
        if form == 'dense' and _is_sparse(A):
            engine.set_matrix(A.toarray(), slot)
        else:
            engine.set_matrix(A, slot)
        self.obj[slot] = A
        self.form[slot] = form
        epoch = getattr(engine, "matrix_epoch", None)
        self.epoch[slot] = None if epoch is None else epoch[slot]
        if slot == 0:
            self.obj[1] = None      # n may have changed; slot 1 is re-uploaded on demand


def _key(cand_id, generation, attempt):
    return ((int(cand_id) & 0xffffffff) << 32) | ((int(generation) & 0xffffff) << 8) | (int(attempt) & 0xff)


def _lu_available(engine, M=None):
    """The direct solve of the ladder (AMS:57 / 59).  A dense matrix: the batched LU up to LU_MAX_N.  A SPARSE matrix of order
    <= LU_MAX_N: its dense form is attached to the resident CSR copy on first need (the reference falls back to SuperLU here;
    same x up to rounding).  Beyond LU_MAX_N there is no device direct solver: the try counts as failed (deviations)."""
    if engine.n > LU_MAX_N:
        return False
    if not engine.is_sparse or getattr(engine, "has_dense_form", False):
        return True
    if M is not None and _is_sparse(M) and hasattr(engine, "add_dense_form"):
        engine.add_dense_form(M.toarray())
        return True
    return False


def _ladder(engine, ptype, v_or_x, lam, stuck, base_psi, max_attempts, pref, matrix_sparse_semantics, cand_id,
            first_method_failed=True, M=None):
    """AMS:43-104 for ONE candidate whose attempt-0 try with the preferred method already failed on the device.
    Returns (x or None, num_psi_attempts)."""
    fallback = 'iterative_gmres' if pref == 'direct_solve' else 'direct_solve'
    method, attempts = pref, 0
    pending_failure = first_method_failed
    engine.upload_vectors(v_or_x[None, :])
    while attempts < max_attempts:
        if not pending_failure:
            psi = psi_magnitude(base_psi, attempts, stuck)
            key = None if matrix_sparse_semantics else [_key(cand_id, engine.generation, attempts + 1)]
            m = _METHOD.get(method)
            if m is None or (m == _abi.METHOD_LU and not _lu_available(engine, M)):
                status = _abi.ST_ZERO_PIVOT      # no direct solver for this operator on the device: the try fails (AMS:98)
                X = None
            else:
                X, st, _ = engine.solve_shifted([lam if ptype == _abi.EIGENVALUE else 0j], [complex(psi).real],
                                                rng_key=key, method=m, use_jacobi=[1 if stuck > 1 else 0],
                                                RHS=None, rhs_shared=(ptype == _abi.SOLVE_LINEAR_SYSTEM))
                status = int(st[0])
            if status == _abi.ST_OK:
                return X[0], attempts
        pending_failure = False
        if method == pref and pref != fallback and attempts == 0:          # AMS:99-102
            method = fallback
            attempts = 0
            continue
        attempts += 1                                                      # AMS:103
    return None, attempts


def _ladder_rs(rs, ptype, v_or_x, lam, stuck, alpha, base_psi, max_attempts, pref, engine=None, M=None, b=None):
    """``_ladder`` on the row-sharded operator (every rank walks it identically: the status words are global).  A GMRES attempt
    is ONE collective call that also mixes and takes the residual (phases solve | mix | residual).  A direct-solve attempt runs
    on a REPLICATED dense copy while the order allows the batched LU (n <= LU_MAX_N: every rank factors redundantly, the matrix
    is small); beyond that there is no device direct solver and the try counts as failed (AMS:57 is SuperLU).
    Returns (vector or None, residual, status, num_psi_attempts)."""
    fallback = 'iterative_gmres' if pref == 'direct_solve' else 'direct_solve'
    method, attempts, pending_failure = pref, 0, True
    eigen = ptype == _abi.EIGENVALUE
    while attempts < max_attempts:
        m = _METHOD.get(method)
        if not pending_failure and m == _abi.METHOD_GMRES:
            psi = psi_magnitude(base_psi, attempts, stuck)
            V1 = np.ascontiguousarray(v_or_x[None, :], dtype=np.complex128).copy()
            out = rs.step(ptype, V1, [alpha], [complex(psi).real], use_jacobi=[1 if stuck > 1 else 0],
                          sigma=[lam] if eigen else None, phases=14)
            st = int(out["status"][0])
            if st in (_abi.ST_OK, _abi.ST_MIX_COLLAPSED):
                return V1[0], float(out["resid"][0]), st, attempts
        elif not pending_failure and m == _abi.METHOD_LU and engine is not None and M is not None and rs.n <= LU_MAX_N:
            psi = psi_magnitude(base_psi, attempts, stuck)
            cache = getattr(engine, "_matrix_cache", None)
            if cache is None:
                cache = engine._matrix_cache = _MatrixCache()
            cache.ensure(engine, M, 0, 'dense')
            if not eigen:
                engine.set_rhs(b)
            engine.upload_vectors(np.ascontiguousarray(v_or_x[None, :], dtype=np.complex128))
            _, st0, _ = engine.solve_shifted([lam if eigen else 0j], [complex(psi).real], rng_key=None, method=_abi.METHOD_LU,
                                             RHS=None, rhs_shared=not eigen, want_x=False)
            if int(st0[0]) == _abi.ST_OK:
                Vn, r1, _, st1 = engine.mix_residual(ptype, [alpha], [lam])
                return Vn[0], float(r1[0]), int(st1[0]), attempts
        pending_failure = False
        if method == pref and pref != fallback and attempts == 0:          # AMS:99-102
            method = fallback
            attempts = 0
            continue
        attempts += 1                                                      # AMS:103
    return None, np.inf, _abi.ST_GMRES_NOCONV, attempts


def step_population(candidates, M, b, strat_params, problem_knowledge, engine, cache=None):
    """Batched equivalent of the loop AMS:574-576.  Returns the number of candidates stepped on the GPU."""
    if not candidates:
        return 0
    State = type(candidates[0]).State
    live = [c for c in candidates if c.state not in (State.CONVERGED, State.RETIRED)]          # AMS:575
    hermitian = bool(problem_knowledge.get('is_hermitian', False))
    gpu, svd, herm = [], [], []
    for c in live:
        pt = c.problem_type.value
        if pt == _abi.SOLVE_LINEAR_SYSTEM or (pt == _abi.EIGENVALUE and not hermitian):
            gpu.append(c)
        elif pt == _abi.SVD and hasattr(engine, "svd_step") and c.problem_matrix is M and not _is_sparse(M):
            svd.append(c)                     # SVD power sweep (AMS:227-255): the first "next" row, SURVEY.md 8f-1
        elif (pt == _abi.EIGENVALUE and hermitian and not _is_sparse(M) and hasattr(engine, "project")
              and isinstance(c.v_k, np.ndarray)):
            herm.append(c)                    # dense Hermitian shortcut with ONE shared eigh
        else:
            # sparse Hermitian shortcut (eigsh, AMS:187-213) and anything else: the reference method, untouched
            c.update_solution_step(M, b, strat_params, problem_knowledge)
    if svd:
        _step_group_svd(svd, M, b, strat_params, engine, State)
    if herm:
        _step_group_hermitian(herm, M, b, strat_params, problem_knowledge, engine, State)
    if not gpu:
        return len(svd) + len(herm)
    cache = cache if cache is not None else getattr(engine, "_matrix_cache", None)
    if cache is None:
        cache = engine._matrix_cache = _MatrixCache()

    ptype = gpu[0].problem_type.value
    N = gpu[0].N_diag
    aggr = strat_params.get('overall_psi_aggression_factor', 1.0)                              # AMS:149-152
    max_retries = strat_params.get('max_psi_retries', MAX_PSI_ATTEMPTS)
    pref = problem_knowledge.get('local_solver_preference', 'direct_solve')
    is_sparse = bool(problem_knowledge.get('is_sparse_problem', False))
    base_psi = PSI_EPSILON_BASE * aggr                                                         # AMS:224
    conv_tol = strat_params.get('current_convergence_threshold', CONVERGENCE_RESIDUAL_TOL)

    # matrix residency.  Direct solves need the dense form; a sparse problem that prefers the direct solver
    # (reference: SuperLU, AMS:57) is densified when it fits the batched LU, otherwise that try fails -> GMRES.
    m_sparse = _is_sparse(M)
    # Row-sharded operator (BASELINE config 5 as worded): when the context was set up with ``enable_row_sharding`` a sparse
    # problem on the GMRES path keeps only n / G rows of the matrix per GPU; every rank steps ALL candidates jointly.
    rs = getattr(engine, "rowshard", None)
    if not (rs is not None and m_sparse and (pref == 'iterative_gmres' or N > LU_MAX_N)
            and all(c.problem_matrix is M for c in gpu)):
        rs = None
    if rs is not None:
        rs.ensure_matrix(M)
        if ptype == _abi.SOLVE_LINEAR_SYSTEM:
            rs.set_rhs(b)
        for c in gpu:
            c.b_vector = b                                                                     # AMS:146
            c.prev_residual = c.residual_k                                                     # AMS:147
        _step_group(gpu, ptype, N, b, engine, State, base_psi, max_retries, pref, is_sparse, conv_tol, _abi.SLOT_CURRENT, rs,
                    M_cur=M)
        return len(gpu)
    if pref == 'direct_solve' and m_sparse and N <= LU_MAX_N:
        form = 'dense'
    else:
        form = 'sparse' if m_sparse else 'dense'
    cache.ensure(engine, M, 0, form)
    if ptype == _abi.SOLVE_LINEAR_SYSTEM:
        engine.set_rhs(b)

    for c in gpu:
        c.b_vector = b                                                                         # AMS:146
        c.prev_residual = c.residual_k                                                         # AMS:147

    # candidates are grouped by the matrix their residual uses (ctor-time matrix, AMS:118, 295)
    groups = {}
    for c in gpu:
        key = 0 if c.problem_matrix is M else id(c.problem_matrix)
        groups.setdefault(key, []).append(c)

    for gkey, cands in groups.items():
        res_slot = _abi.SLOT_CURRENT
        if gkey != 0:
            cache.ensure(engine, cands[0].problem_matrix, 1, 'sparse' if _is_sparse(cands[0].problem_matrix) else 'dense')
            res_slot = _abi.SLOT_CTOR
        _step_group(cands, ptype, N, b, engine, State, base_psi, max_retries, pref, is_sparse, conv_tol, res_slot, M_cur=M)
    return len(gpu)


def _step_group(cands, ptype, N, b, engine, State, base_psi, max_retries, pref, is_sparse, conv_tol, res_slot, rs=None,
                M_cur=None):
    C_ = len(cands)
    eigen = ptype == _abi.EIGENVALUE
    # the vectors travel through one page-locked staging buffer (H2D before, D2H after the fused step)
    V = engine.staging((C_, N)) if hasattr(engine, "staging") else np.empty((C_, N), dtype=np.complex128)
    for i, c in enumerate(cands):
        if eigen:
            if np.linalg.norm(c.v_k) < 1e-10:                                                  # AMS:259-263
                c.v_k = (np.random.rand(N) + 1j * np.random.rand(N))
                c.v_k /= np.linalg.norm(c.v_k)
                c.stuck_counter += 1
                c.num_resets += 1
                print(f"Candidate {c.id}: Eigenvector collapsed, reinitialized randomly before InverseIterateSolver.")
            V[i] = c.v_k
        else:
            V[i] = c.x_k
    V_before = V.copy()
    stuck = np.array([c.stuck_counter for c in cands], dtype=np.int64)
    alpha = np.array([complex(c.alpha_local_step).real for c in cands], dtype=np.float64)
    psi0 = np.array([complex(psi_magnitude(base_psi, 0, int(s))).real for s in stuck], dtype=np.float64)
    gen = engine.generation
    keys = None if is_sparse else np.array([_key(c.id, gen, 0) for c in cands], dtype=np.uint64)
    method0 = _METHOD.get(pref)
    if rs is not None and method0 == _abi.METHOD_GMRES:
        out = rs.step(ptype, V, alpha, psi0, use_jacobi=(stuck > 1).astype(np.uint8), phases=15)
    elif method0 is None or (method0 == _abi.METHOD_LU and (rs is not None or engine.is_sparse or N > LU_MAX_N)):
        # unknown method / direct solve beyond the LU limit (sparse or dense): the first try fails for everybody (AMS:92, 98)
        lam0 = np.zeros(C_, dtype=np.complex128)
        if eigen:
            lam0 = rs.step(ptype, V, phases=1)["lam"] if rs is not None else engine.rq(V)[0]
        out = dict(lam=lam0,
                   resid=np.full(C_, np.inf), mixnorm=np.zeros(C_), status=np.full(C_, _abi.ST_ZERO_PIVOT, dtype=np.int32),
                   iters=np.zeros(C_, dtype=np.int32))
    else:
        out = engine.step(ptype, alpha, psi0, V=V, rng_key=keys, method=method0,
                          use_jacobi=(stuck > 1).astype(np.uint8), res_slot=res_slot)
    lam, resid, status = out["lam"], out["resid"], out["status"]

    for i, c in enumerate(cands):
        st = int(status[i])
        # MIX_COLLAPSED = the solve succeeded but ||(1-a)v + a x|| <= 1e-10 (AMS:283): still the success branch
        solved, retries = (st in (_abi.ST_OK, _abi.ST_MIX_COLLAPSED)), 0
        if eigen:
            c.lambda_k = np.complex128(lam[i])                                                  # AMS:268 (0 when |<v,v>| tiny)
        need_residual = False
        if st in (_abi.ST_ZERO_PIVOT, _abi.ST_NONFINITE, _abi.ST_GMRES_NOCONV) and rs is not None:
            vec, r1, st1, retries = _ladder_rs(rs, ptype, V_before[i], complex(lam[i]), int(stuck[i]), float(alpha[i]), base_psi,
                                               max_retries, pref, engine, M_cur, b)
            if vec is not None:
                V[i] = vec
                resid[i] = r1
                st = st1
                solved = True
        elif st in (_abi.ST_ZERO_PIVOT, _abi.ST_NONFINITE, _abi.ST_GMRES_NOCONV):
            # the preferred method failed at attempt 0: walk the rest of the ladder for this candidate
            x, retries = _ladder(engine, ptype, V_before[i], complex(lam[i]), int(stuck[i]), base_psi, max_retries,
                                 pref, is_sparse, c.id, M=M_cur)
            if x is not None:
                Vn, r1, mixn, st1 = engine.mix_residual(ptype, [alpha[i]], [lam[i]], res_slot=res_slot)
                V[i] = Vn[0]
                resid[i] = r1[0]
                st = int(st1[0])
                solved = True
        if solved:
            c.local_psi_retries_needed = retries                                                # AMS:278
            if eigen:
                if st == _abi.ST_MIX_COLLAPSED:                                                 # AMS:283
                    c.v_k = (np.random.rand(N) + 1j * np.random.rand(N)) / np.sqrt(N)
                    need_residual = True
                else:
                    c.v_k = V[i].copy()                                                         # AMS:280-282
            else:
                c.x_k = V[i].copy()                                                             # AMS:285
            c.stuck_counter = max(0, c.stuck_counter - 1)                                       # AMS:286
        else:
            # RuntimeError branch, AMS:287-293 (V_COLLAPSED cannot happen: the guard above ran on the host)
            c.stuck_counter += 1
            c.w_k *= 0.001
            c.alpha_local_step = max(c.alpha_local_step * 0.5, 1e-6)
            if c.stuck_counter >= MAX_STUCK_FOR_RETIREMENT:
                c.state = State.RETIRED
                c.num_resets += 1
            else:
                c.state = State.STUCK
                c.initialize_random_solution()
            need_residual = True
        if need_residual:                                                                       # AMS:295-299
            vec = c.v_k if eigen else c.x_k
            if rs is not None:
                V1 = np.ascontiguousarray(vec[None, :], dtype=np.complex128).copy()
                resid[i] = rs.step(ptype, V1, sigma=[c.lambda_k] if eigen else None, phases=8)["resid"][0]
            else:
                resid[i] = engine.residual(ptype, vec[None, :], [c.lambda_k] if eigen else None, res_slot=res_slot)[0]
        c.residual_k = np.float64(resid[i])
        c.param_history.append(c.get_current_solution_params())                                # AMS:303-304
        c.residual_history.append(c.residual_k)
        _adapt_and_test(c, State, conv_tol)


def _step_group_svd(cands, M, b, strat_params, engine, State):
    """SVD branch of update_solution_step for a batch (AMS:146-147, 227-255, 300-304, 306-331); the inverse-iteration
    solver is never called on this branch (the reference only constructs it, AMS:224)."""
    cache = getattr(engine, "_svd_cache", None)
    if cache is not M:
        engine.svd_set_matrix(M)
        engine._svd_cache = M
    Mr, Mc = M.shape
    conv_tol = strat_params.get('current_convergence_threshold', CONVERGENCE_RESIDUAL_TOL)
    for c in cands:
        c.b_vector = b
        c.prev_residual = c.residual_k
    U = np.ascontiguousarray(np.stack([c.u_k for c in cands]), dtype=np.complex128)
    V = np.ascontiguousarray(np.stack([c.right_v_k for c in cands]), dtype=np.complex128)
    out = engine.svd_step(U, V)
    redo = []
    for i, c in enumerate(cands):
        st = int(out["status"][i])
        failed = False
        if st == _abi.ST_V_COLLAPSED:                                                           # AMS:229-232
            c.right_v_k = (np.random.rand(Mc) + 1j * np.random.rand(Mc))
            c.right_v_k /= np.linalg.norm(c.right_v_k)
            c.stuck_counter += 1
            c.num_resets += 1
            failed = True
        elif st == _abi.ST_MIX_COLLAPSED:                                                       # AMS:236-239
            c.u_k = (np.random.rand(Mr) + 1j * np.random.rand(Mr))
            c.u_k /= np.linalg.norm(c.u_k)
            c.stuck_counter += 1
            c.num_resets += 1
            failed = True
        else:
            c.sigma_k = np.float64(out["sigma"][i])                                             # AMS:234, 241
            c.u_k = U[i].copy()
            c.right_v_k = V[i].copy()
            if c.sigma_k < SIGMA_SIMILARITY_TOL_ABS / 100:                                      # AMS:243-247
                c.state = State.CONVERGED
                c.stuck_counter = 0
                replaced = False
                if np.linalg.norm(c.u_k) < 1e-10:                                               # AMS:246
                    c.u_k = np.ones(Mr, dtype=np.complex128) / np.sqrt(Mr)
                    replaced = True
                if np.linalg.norm(c.right_v_k) < 1e-10:                                         # AMS:247
                    c.right_v_k = np.ones(Mc, dtype=np.complex128) / np.sqrt(Mc)
                    replaced = True
                if replaced:
                    redo.append(c)              # the device residual was taken on the un-replaced vectors (AMS:301 uses the new ones)
            else:
                c.stuck_counter = max(0, c.stuck_counter - 1)                                   # AMS:248
            if not (redo and redo[-1] is c):
                c.residual_k = np.float64(out["resid"][i])                                      # AMS:301
        if failed:                                                                              # AMS:249-255
            c.stuck_counter += 1
            c.w_k *= 0.001
            c.alpha_local_step *= 0.5
            c.state = State.STUCK
            if c.stuck_counter >= MAX_STUCK_FOR_RETIREMENT:
                c.state = State.RETIRED
            c.u_k = (np.random.rand(Mr) + 1j * np.random.rand(Mr)) / np.sqrt(Mr)
            c.right_v_k = (np.random.rand(Mc) + 1j * np.random.rand(Mc)) / np.sqrt(Mc)
            c.sigma_k = 1.0
            redo.append(c)
    if redo:
        r = engine.svd_residual(np.stack([c.u_k for c in redo]), np.stack([c.right_v_k for c in redo]),
                                [float(c.sigma_k) for c in redo])
        for c, ri in zip(redo, r):
            c.residual_k = np.float64(ri)
    for c in cands:
        c.param_history.append(c.get_current_solution_params())                                # AMS:303-304
        c.residual_history.append(c.residual_k)
        _adapt_and_test(c, State, conv_tol)


def _adapt_and_test(c, State, conv_tol):
    """AMS:306-331, scalar host logic kept bit-identical to the reference."""
    if c.prev_residual > 1e-10:
        if c.residual_k < c.prev_residual * 0.9:
            c.alpha_local_step = min(c.alpha_local_step * 1.1, 1.0)
            if c.state != State.CONVERGED:
                c.state = State.REFINING
        elif c.residual_k > c.prev_residual * 1.5 and c.prev_residual > 1e-5:
            c.alpha_local_step = max(c.alpha_local_step * 0.5, 1e-6)
            if c.state != State.CONVERGED:
                c.state = State.STUCK
        else:
            c.alpha_local_step = max(c.alpha_local_step * 0.95, 1e-6)
            if c.state not in [State.CONVERGED, State.STUCK, State.RETIRED]:
                c.state = State.EXPLORING
    params = c.get_current_solution_params()
    finite = False
    if params is not None:
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
    if c.residual_k < conv_tol and finite:
        c.state = State.CONVERGED
        c.w_k = 1.0
        c.stuck_counter = 0
        c.alpha_local_step = 0.0


# ---------------------------------------------------------------------------------------------------------------
# drop-in installation
# ---------------------------------------------------------------------------------------------------------------
def install_dropin(ams_module, engine):
    """Seam A: rebind the module-global solver name the reference looks up on every step (AMS:224)."""
    GpuInverseIterateSolver.bind_engine(engine)
    ams_module.InverseIterateSolver = GpuInverseIterateSolver
    return ams_module


def _step_group_hermitian(cands, M, b, strat_params, problem_knowledge, engine, State):
    """Dense Hermitian shortcut (AMS:155-186) for a whole group.  The reference calls ``sla.eigh(current_matrix_A)`` inside
    EVERY candidate's step -- C identical O(n^3) factorizations per generation.  Here ONE eigendecomposition per matrix runs
    on the device (``engine.heev``: cyclic Jacobi, heev.cu; same eigenpairs to rounding, vectors up to a phase), the
    similarity scores |v_k^H E| of all candidates are one device GEMM (``engine.project``) and the residuals one batched
    device pass.  A failing decomposition falls back to the reference method per candidate (AMS:182-185)."""
    import scipy.linalg as sla
    cache = getattr(engine, "_eigh_cache", None)
    try:
        if cache is None or cache[0] is not M:
            if hasattr(engine, "heev"):
                w, E = engine.heev(M)                                               # AMS:161 on the device (Jacobi, heev.cu)
            else:
                w, E = sla.eigh(M)                                                  # engines without it (CPU test stand-in)
            cache = engine._eigh_cache = (M, w, E, np.ascontiguousarray(E.conj()))  # conj(E) C-order = E^H column-major
    except Exception:                                                              # AMS:182-185: "Falling back."
        for c in cands:
            c.update_solution_step(M, b, strat_params, problem_knowledge)
        return
    _, w, E, Ec = cache
    n = E.shape[0]
    usable = [c for c in cands if c.v_k.shape[0] == n and E.shape[1] > 0]
    for c in cands:
        if c not in usable:
            c.update_solution_step(M, b, strat_params, problem_knowledge)
    if not usable:
        return
    V = np.ascontiguousarray(np.stack([c.v_k for c in usable]), dtype=np.complex128)
    scores = np.abs(engine.project(Ec, V))                                          # |E^H v| = |v^H E| (AMS:165), [C][n]
    best = np.argmax(scores, axis=1)                                                # AMS:169
    Vn = np.empty_like(V)
    lam = np.empty(len(usable), dtype=np.complex128)
    for k, c in enumerate(usable):
        c.b_vector = b                                                              # AMS:146
        c.prev_residual = c.residual_k                                              # AMS:147
        v = E[:, best[k]].astype(np.complex128, copy=True)
        v /= np.linalg.norm(v)                                                      # AMS:173
        Vn[k] = v
        lam[k] = w[best[k]]
    cache_m = getattr(engine, "_matrix_cache", None)
    if cache_m is None:
        cache_m = engine._matrix_cache = _MatrixCache()
    cache_m.ensure(engine, M, 0, 'dense')
    resid = engine.residual(_abi.EIGENVALUE, Vn, lam)                               # AMS:175, batched
    for k, c in enumerate(usable):
        c.lambda_k = w[best[k]]                                                     # a real float64, like the reference (AMS:171)
        c.v_k = Vn[k].copy()
        c.residual_k = np.float64(resid[k])
        c.state = State.CONVERGED                                                   # AMS:176-179
        c.stuck_counter = 0
        c.local_psi_retries_needed = 0
        c.w_k = 1.0
        c.param_history.append(c.get_current_solution_params())                     # AMS:216-217
        c.residual_history.append(c.residual_k)


def gpu_generation(maus_solver, iteration, engine):
    """One generation of ``MAUS_Solver.evolve`` (the body of the loop at AMS:572-577) with the candidate loop
    executed by ``step_population``.  ``evolve`` itself cannot be used: it reads an undefined global (AMS:583)."""
    maus_solver._update_global_diagnostics(iteration)
    maus_solver._adjust_global_strategy(iteration)
    n = step_population(maus_solver.candidates, maus_solver.M, maus_solver.b, maus_solver.strat_params,
                        maus_solver.problem_knowledge, engine)
    maus_solver._manage_candidates(iteration)
    return n

"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REAL reference.

Run in the build container (needs /root/reference):   python oracle/gen_golden.py
Every recorded step re-seeds the global RNGs (``np.random.seed(s); random.seed(s)``) right before the
reference's ``update_solution_step`` so a checker can replay a single step in isolation.  Nothing in the
reference file is edited (see oracle/ref_loader.py for the two outside work-arounds).

Scenarios:
  eig8        AMS:654-657 scenario 2A: eigen, N=8 general complex, 30 candidates, dense direct
  eig100      K1 hot-path variant: create_laplace_like_complex_eigen_for_MAUS(100, False), 16 candidates, tol 1e-10
  lin5_shim / lin5_shipped   AMS:644-653 scenario 1 (dynamic Ax=b, N=5) with / without the gmres tol->rtol shim
  gmres64     dense ill-conditioned Ax=b, n=64, 'Fragile' -> GMRES preferred, half the candidates stuck=2 (Jacobi)
  speig200    sparse CSC eigenproblem n=200 -> 'Critical' -> GMRES + sparse Psi
  fail6       non-finite operator: every attempt fails -> RuntimeError branch AMS:287-293
"""
import json
import os
import random
import sys

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_loader import load_reference, quiet  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 20260


def snap(c):
    al = c.alpha_local_step
    return dict(
        lam=complex(c.lambda_k) if c.lambda_k is not None else complex("nan"),
        v=None if c.v_k is None else np.array(c.v_k, dtype=np.complex128, copy=True),
        u=None if getattr(c, "u_k", None) is None else np.array(c.u_k, dtype=np.complex128, copy=True),
        rv=None if getattr(c, "right_v_k", None) is None else np.array(c.right_v_k, dtype=np.complex128, copy=True),
        sigma=float(np.real(c.sigma_k)) if getattr(c, "sigma_k", None) is not None else float("nan"),
        x=None if c.x_k is None else np.array(c.x_k, dtype=np.complex128, copy=True),
        state=int(c.state.value), w=float(c.w_k), res=float(c.residual_k), prev=float(c.prev_residual),
        alpha=complex(al), alpha_is_complex=bool(isinstance(al, (complex, np.complexfloating))),
        stuck=int(c.stuck_counter), retries=int(c.local_psi_retries_needed), resets=int(c.num_resets),
        hist=len(c.residual_history),
    )


class Recorder:
    def __init__(self, keep):
        self.keep = keep
        self.records = []
        self.counter = 0

    def __call__(self, cand, solver):
        idx = self.counter
        self.counter += 1
        if not self.keep(idx, solver):
            cand.update_solution_step(solver.M, solver.b, solver.strat_params, solver.problem_knowledge)
            return
        before = snap(cand)
        seed = SEED * 1000 + idx
        np.random.seed(seed % (2 ** 32))
        random.seed(seed)
        strat = dict(solver.strat_params)
        know = dict(solver.problem_knowledge)
        cand.update_solution_step(solver.M, solver.b, solver.strat_params, solver.problem_knowledge)
        after = snap(cand)
        ctor_is_current = cand.problem_matrix is solver.M
        self.records.append(dict(seed=seed, cand_id=int(cand.id), before=before, after=after,
                                 strat={k: float(v) for k, v in strat.items()},
                                 pref=know["local_solver_preference"], sparse=bool(know["is_sparse_problem"]),
                                 ctor_is_current=bool(ctor_is_current)))


def _pack(vecs, n):
    out = np.full((len(vecs), n), np.nan + 1j * np.nan, dtype=np.complex128)
    for i, v in enumerate(vecs):
        if v is not None:
            out[i] = v
    return out


def save(name, records, A, b, A_ctor, meta):
    n = A.shape[0] if meta.get("problem_type") != 3 else max(A.shape)
    arrs = {}
    if sp.issparse(A):
        Ac = sp.csc_matrix(A)
        arrs.update(A_indptr=Ac.indptr.astype(np.int64), A_indices=Ac.indices.astype(np.int64), A_data=Ac.data)
        meta["A_format"] = "csc"
    else:
        arrs["A"] = np.asarray(A, dtype=np.complex128)
        meta["A_format"] = "dense"
    if b is not None:
        arrs["b"] = np.asarray(b, dtype=np.complex128)
    if A_ctor is not None:
        arrs["A_ctor"] = np.asarray(A_ctor, dtype=np.complex128)
    for side in ("before", "after"):
        arrs[f"{side}_v"] = _pack([r[side]["v"] for r in records], n)
        arrs[f"{side}_x"] = _pack([r[side]["x"] for r in records], n)
        if meta.get("problem_type") == 3:
            arrs[f"{side}_u"] = _pack([r[side]["u"] for r in records], A.shape[0])
            arrs[f"{side}_rv"] = _pack([r[side]["rv"] for r in records], A.shape[1])
            arrs[f"{side}_sigma"] = np.array([r[side]["sigma"] for r in records], dtype=np.float64)
        arrs[f"{side}_lam"] = np.array([r[side]["lam"] for r in records], dtype=np.complex128)
        arrs[f"{side}_alpha"] = np.array([r[side]["alpha"] for r in records], dtype=np.complex128)
        for k in ("w", "res", "prev"):
            arrs[f"{side}_{k}"] = np.array([r[side][k] for r in records], dtype=np.float64)
        for k in ("state", "stuck", "retries", "resets", "hist"):
            arrs[f"{side}_{k}"] = np.array([r[side][k] for r in records], dtype=np.int64)
        arrs[f"{side}_alpha_is_complex"] = np.array([r[side]["alpha_is_complex"] for r in records], dtype=np.bool_)
    arrs["seed"] = np.array([r["seed"] for r in records], dtype=np.int64)
    arrs["cand_id"] = np.array([r["cand_id"] for r in records], dtype=np.int64)
    arrs["ctor_is_current"] = np.array([r["ctor_is_current"] for r in records], dtype=np.bool_)
    meta["steps"] = [dict(strat=r["strat"], pref=r["pref"], sparse=r["sparse"]) for r in records]
    arrs["meta"] = np.array(json.dumps(meta))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {len(records)} steps, {os.path.getsize(path) / 1024:.0f} KiB")


def run_maus(ams, solver, gens, keep):
    rec = Recorder(keep)
    for i in range(gens):
        from ref_loader import drive_generation
        quiet(drive_generation, solver, i + 1, rec)
    return rec.records


def versions():
    import scipy
    return dict(numpy=np.__version__, scipy=scipy.__version__, seed=SEED)


def main():
    shim = load_reference(gmres_shim=True, name="ams_shim")
    shipped = load_reference(gmres_shim=False, name="ams_shipped")

    # eig8 -- AMS:654-657
    np.random.seed(SEED); random.seed(SEED)
    M = shim.create_laplace_like_complex_eigen_for_MAUS(8, make_hermitian=False)
    s = quiet(shim.MAUS_Solver, M, problem_type=shim.ProblemType.EIGENVALUE, initial_num_candidates=30,
              global_convergence_tol=1e-7)
    recs = run_maus(shim, s, 12, lambda i, sv: True)
    save("eig8", recs, s.M, None, None, dict(problem_type=1, gmres_mode="shim", **versions()))

    # eig100 -- K1 hot-path variant
    np.random.seed(SEED + 1); random.seed(SEED + 1)
    M = shim.create_laplace_like_complex_eigen_for_MAUS(100, make_hermitian=False)
    s = quiet(shim.MAUS_Solver, M, problem_type=shim.ProblemType.EIGENVALUE, initial_num_candidates=16,
              global_convergence_tol=1e-10)
    recs = run_maus(shim, s, 20, lambda i, sv: (i % 13) == 0 or i < 16)
    save("eig100", recs, s.M, None, None, dict(problem_type=1, gmres_mode="shim", **versions()))

    # lin5 -- AMS:644-653, both gmres modes
    for tag, ams in (("lin5_shim", shim), ("lin5_shipped", shipped)):
        np.random.seed(SEED + 2); random.seed(SEED + 2)
        s = quiet(ams.MAUS_Solver, np.eye(5), problem_type=ams.ProblemType.SOLVE_LINEAR_SYSTEM,
                  b_vector=np.ones(5), initial_num_candidates=15, global_convergence_tol=1e-7)
        A_ctor = s.candidates[0].problem_matrix
        A_final, b_final = ams.create_dynamic_solve_matrix_and_b(N=5, t_step=19, time_max_iter=20)
        s.M = A_final; s.b = b_final
        s.diag_info = s._diagnose_matrix_initial(A_final)
        s.problem_knowledge.update({'is_hermitian': s.diag_info.get('is_hermitian', False),
                                    'is_complex_symmetric': s.diag_info.get('is_complex_symmetric', False),
                                    'is_sparse_problem': s.diag_info.get('is_sparse_init', False)})
        recs = run_maus(ams, s, 8, lambda i, sv: True)
        A_ctor_dense = A_ctor.toarray() if sp.issparse(A_ctor) else np.asarray(A_ctor)
        save(tag, recs, s.M, s.b, A_ctor_dense,
             dict(problem_type=2, gmres_mode="shim" if ams is shim else "as_shipped", **versions()))

    # gmres64 -- dense ill-conditioned linear system, GMRES preferred, Jacobi on for forced-stuck candidates
    np.random.seed(SEED + 3); random.seed(SEED + 3)
    n = 64
    A = sla.hilbert(n).astype(np.complex128) + 1e-9 * np.eye(n) \
        + 1e-3 * np.diag(np.linspace(1, 2, n) + 1j * np.linspace(-1, 1, n))
    b = A @ np.ones(n, dtype=np.complex128)
    s = quiet(shim.MAUS_Solver, A, problem_type=shim.ProblemType.SOLVE_LINEAR_SYSTEM, b_vector=b,
              initial_num_candidates=8, global_convergence_tol=1e-8)
    for k, c in enumerate(s.candidates):
        if k % 2 == 1:
            c.stuck_counter = 3
    recs = run_maus(shim, s, 4, lambda i, sv: True)
    save("gmres64", recs, s.M, s.b, None, dict(problem_type=2, gmres_mode="shim", cond=float(s.cond_number),
                                                  **versions()))

    # speig200 -- sparse CSC eigenproblem
    np.random.seed(SEED + 4); random.seed(SEED + 4)
    n = 200
    rng = np.random.default_rng(SEED + 4)
    S = sp.random(n, n, density=0.05, random_state=rng, format="csc", dtype=np.float64)
    S = S + 1j * sp.random(n, n, density=0.05, random_state=rng, format="csc", dtype=np.float64)
    S = sp.csc_matrix(S + sp.diags(4.0 + np.linspace(0, 3, n) + 1j * np.linspace(-1, 1, n), format="csc"))
    s = quiet(shim.MAUS_Solver, S, problem_type=shim.ProblemType.EIGENVALUE, initial_num_candidates=6,
              global_convergence_tol=1e-8)
    for k, c in enumerate(s.candidates):
        if k % 3 == 2:
            c.stuck_counter = 2
    recs = run_maus(shim, s, 5, lambda i, sv: (i % 3) == 0)
    save("speig200", recs, s.M, None, None, dict(problem_type=1, gmres_mode="shim", **versions()))

    # fail6 -- non-finite operator drives the failure branch
    np.random.seed(SEED + 5); random.seed(SEED + 5)
    n = 6
    A = (np.random.rand(n, n) + 1j * np.random.rand(n, n)).astype(np.complex128)
    s = quiet(shim.MAUS_Solver, A, problem_type=shim.ProblemType.EIGENVALUE, initial_num_candidates=3,
              global_convergence_tol=1e-8)
    bad = A.copy(); bad[2, 3] = np.inf
    s.M = bad
    s.strat_params["max_psi_retries"] = 3
    recs = run_maus(shim, s, 3, lambda i, sv: True)
    save("fail6", recs, s.M, None, A, dict(problem_type=1, gmres_mode="shim", **versions()))


def svd_golden():
    """svd5x4 -- AMS:662-665 scenario 3 (near-low-rank 5 x 4) and a 40 x 28 variant; SVD power-sweep branch."""
    shim = load_reference(gmres_shim=True, name="ams_svd")
    for tag, (mr, mc, ncand, gens) in {"svd5x4": (5, 4, 25, 12), "svd40x28": (40, 28, 12, 6)}.items():
        np.random.seed(SEED + 6 + mr); random.seed(SEED + 6 + mr)
        M = shim.create_low_rank_svd_matrix_for_MAUS(mr, mc, target_rank=2)
        s = quiet(shim.MAUS_Solver, M, problem_type=shim.ProblemType.SVD, initial_num_candidates=ncand,
                  global_convergence_tol=1e-6)
        recs = run_maus(shim, s, gens, (lambda i, sv: True) if mr < 10 else (lambda i, sv: i % 5 == 0))
        save(tag, recs, s.M, None, None, dict(problem_type=3, gmres_mode="shim", **versions()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "svd":
        svd_golden()
    else:
        main()
        svd_golden()

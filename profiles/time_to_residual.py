"""time-to-residual 1e-10 (second half of BASELINE.json's metric): GPU population vs the reference's CPU algorithm.

GPU : K3 workload (n = 4096, 128 candidates, dense non-Hermitian, direct path, tol 1e-10) through step_population on host
      candidate objects with the full alpha / state / convergence logic, until the first candidate and until 8 DISTINCT
      eigenpairs have residual < 1e-10.
CPU : the oracle (= the reference's numpy/scipy algorithm) stepping ONLY the candidate that converged first on the GPU,
      from the same initial vector, until its residual < 1e-10 (a bounded sample: the reference itself would step all 128
      candidates every generation, so its time-to-first-residual is generations x 128 x the measured step time).
Usage: python profiles/time_to_residual.py [n] [C]"""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adaptive_matrix_solver_b200 as pkg                                    # noqa: E402
from adaptive_matrix_solver_b200.workloads import k2_matrix, initial_vectors  # noqa: E402
from oracle import maus_oracle as mo                                          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
C_ = int(sys.argv[2]) if len(sys.argv) > 2 else 128
TOL = 1e-10
A = k2_matrix(n)
V0 = initial_vectors(C_, n)
strat = dict(overall_psi_aggression_factor=1.0, max_psi_retries=25, current_convergence_threshold=TOL)
know = dict(local_solver_preference="direct_solve", is_sparse_problem=False, is_hermitian=False)

eng = pkg.MausEngine(0)
np.random.seed(1)
cands = [pkg.Candidate(A, pkg.ProblemType.EIGENVALUE, n, initial_lambda=0j, initial_v=V0[i].copy()) for i in range(C_)]
pkg.step_population(cands[:2], A, None, dict(strat, current_convergence_threshold=0.0), know, eng)   # warm-up (allocations)
cands = [pkg.Candidate(A, pkg.ProblemType.EIGENVALUE, n, initial_lambda=0j, initial_v=V0[i].copy()) for i in range(C_)]
t0 = time.perf_counter()
first = None; t_first = None; t_eight = None; gens = 0
State = pkg.Candidate.State
while gens < 200:
    pkg.step_population(cands, A, None, strat, know, eng)
    gens += 1
    conv = [c for c in cands if c.state == State.CONVERGED]
    if conv and first is None:
        first = min(conv, key=lambda c: c.residual_k); t_first = time.perf_counter() - t0; g_first = gens
    distinct = []
    for c in conv:
        if all(abs(c.lambda_k - d.lambda_k) > 1e-5 + 1e-6 * abs(d.lambda_k) or abs(np.vdot(c.v_k, d.v_k)) <= 0.999 for d in distinct):
            distinct.append(c)                                                # AMS:435-436 similarity rule
    if len(distinct) >= 8:
        t_eight = time.perf_counter() - t0
        break
ev = np.linalg.eigvals(A) if n <= 4096 else None
err = max(np.abs(ev - c.lambda_k).min() for c in distinct) if ev is not None else None
idx = cands.index(first)

# CPU: the oracle on the first-converged candidate
o = mo.CandState(problem_type=mo.EIGENVALUE, N=n); o.v_k = V0[idx].copy(); o.lambda_k = 0j
np.random.seed(1)
steps = 0; t1 = time.perf_counter()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    while o.state != mo.CONVERGED and steps < 200:
        mo.candidate_step(o, A, None, strat, know)
        steps += 1
t_cpu = time.perf_counter() - t1
print(json.dumps({
    "workload": f"K3 n={n}, {C_} candidates, tol {TOL}", "gpu_generations_to_first": g_first, "gpu_s_to_first": round(t_first, 3),
    "gpu_generations_to_8_distinct": gens, "gpu_s_to_8_distinct": None if t_eight is None else round(t_eight, 3),
    "max_eig_error_of_converged": err, "first_residual": float(first.residual_k),
    "cpu_candidate_steps_to_1e-10": steps, "cpu_s_single_candidate": round(t_cpu, 2), "cpu_s_per_step": round(t_cpu / max(steps, 1), 3),
    "cpu_final_residual": float(o.residual_k), "cpu_lambda_minus_gpu_lambda": abs(complex(o.lambda_k) - complex(first.lambda_k)),
    "cpu_s_to_first_extrapolated_population": round(t_cpu / max(steps, 1) * C_ * g_first, 1),
    "cpu_cores": os.cpu_count()}), flush=True)
eng.close()

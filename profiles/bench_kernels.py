"""Roofline measurements of the HBM-bound kernels and of the K2/K4/K5 configurations (BASELINE.json configs 2, 4, 5).
Run on a B200:  python profiles/bench_kernels.py [section ...]   -> JSON lines (copy into profiles/).
Timing: CUDA events recorded by the library around each matvec launch (maus_profile_*), after warm-up."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaptive_matrix_solver_b200 as pkg                      # noqa: E402
from adaptive_matrix_solver_b200 import _abi                   # noqa: E402
from adaptive_matrix_solver_b200.workloads import k2_matrix, k4_system, k5_sparse, initial_vectors   # noqa: E402

HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
sections = set(sys.argv[1:]) or {"gemv", "spmm", "vec", "k2", "k4", "k5"}
eng = pkg.MausEngine(0)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def matvec_roofline(tag, C_, reps, alg_bytes):
    V = initial_vectors(C_, eng.n)
    eng.upload_vectors(V)
    for _ in range(3):
        eng.rq(C_=C_)
    eng.profile_reset(True)
    for _ in range(reps):
        eng.rq(C_=C_)
    p = eng.profile_read(); eng.profile_reset(False)
    ms = p["matvec_ms"] / max(1, p["matvec_launches"])
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    emit(kernel=tag, n=eng.n, candidates=C_, launches=p["matvec_launches"], ms_per_launch=round(ms, 4),
         algorithmic_bytes_per_launch=alg_bytes, achieved_gbs=round(gbs, 1), hbm_peak_gbs=HBM, frac=round(gbs / HBM, 3))


if "gemv" in sections:
    for n in (4096, 8192):
        eng.set_matrix(k2_matrix(n))
        for C_ in (1, 4):
            matvec_roofline("gemv_rowmajor_kernel (RQ matvec)", C_, 20, 16.0 * n * n + 32.0 * n * C_)

if "spmm" in sections:
    n = 1_000_000
    A = k5_sparse(n)
    eng.set_matrix(A)
    for C_ in (1, 4, 8):       # 8: one pass through the 128-byte-per-entry layout (csr_spmm_packed8_kernel)
        matvec_roofline("csr_spmm_kernel", C_, 20, 20.0 * A.nnz + 8.0 * (n + 1) + 32.0 * n * C_)

if "vec" in sections:
    # fused vector kernels on long vectors (n = 1M, 8 candidates: the K5 shapes), multi-block reductions; algorithmic bytes =
    # 32 n C (RQ: v, y), 80 n C (mix: read v, x, write v; normalise: read v, write v -- the norm is only known after
    # the first pass and 16 MB vectors do not stay on chip), 32 n C (residual: v, y)
    n = 1_000_000
    if eng.n != n or not eng.is_sparse:
        eng.set_matrix(k5_sparse(n))
    for C_ in (8, 32):
        rng = np.random.default_rng(2)
        V = rng.random((C_, n)) + 1j * rng.random((C_, n)); V /= np.linalg.norm(V, axis=1, keepdims=True)
        eng.upload_vectors(V)
        lam, _ = eng.rq(C_=C_)
        for name, fn, alg in (("rq_part + rq_final (Rayleigh quotient dots)", lambda: eng.rq(C_=C_), 32.0 * n * C_),
                              ("res_part + res_final (residual norm, one pass)", lambda: eng.residual(_abi.EIGENVALUE, lam=lam, C_=C_), 32.0 * n * C_),
                              ("mix_part + mix_apply (mix + normalise)",
                               lambda: eng.mix_residual(_abi.EIGENVALUE, np.full(C_, 0.5), lambda_old=lam, want_v=False), 80.0 * n * C_)):
            for _ in range(2):
                fn()
            eng.profile_reset(True)
            reps = 10
            for _ in range(reps):
                fn()
            bd = eng.profile_breakdown()["vec"]; eng.profile_reset(False)
            calls = reps * (2 if name.startswith("mix") else 1)       # mix_residual also runs the residual family
            ms = bd["ms"] / reps
            byts = bd["work"] / reps
            emit(kernel=name, n=n, candidates=C_, ms_per_call=round(ms, 4), algorithmic_bytes_per_call=byts,
                 achieved_gbs=round(byts / (ms * 1e-3) / 1e9, 1), hbm_peak_gbs=HBM, frac=round(byts / (ms * 1e-3) / 1e9 / HBM, 3),
                 note="mix_residual = mix + residual families together" if name.startswith("mix") else None)

if "k2" in sections:
    # config 2: dense non-Hermitian n=1024, 64 candidates, Psi on, 1 GPU: fused generations, resident
    n, C_ = 1024, 64
    eng.set_matrix(k2_matrix(n))
    eng.upload_vectors(initial_vectors(C_, n))
    alpha = np.full(C_, 0.01); psi = np.full(C_, 1e-20)
    for s in range(3):
        eng.step(_abi.EIGENVALUE, alpha, psi, rng_key=np.arange(C_, dtype=np.uint64))
    t0 = time.perf_counter(); K = 10
    for s in range(K):
        out = eng.step(_abi.EIGENVALUE, alpha, psi, rng_key=np.arange(C_, dtype=np.uint64) + np.uint64(s * 1000))
    dt = time.perf_counter() - t0
    emit(config="K2 dense n=1024 c64 direct", steps=K, ms_per_generation=round(dt / K * 1e3, 3),
         candidate_steps_per_s=round(C_ * K / dt, 1), frac_fp64_peak=round(8 / 3 * n ** 3 * C_ * K / dt / 37.1e12, 4))

if "k4" in sections:
    # config 4: ill-conditioned dense Ax=b n=8192, 128 candidates, Jacobi-preconditioned GMRES for half of them
    n, C_ = 8192, 128
    A, b = k4_system(n)
    eng.set_matrix(A); eng.set_rhs(b)
    rng = np.random.default_rng(0)
    X0 = rng.standard_normal((C_, n)) + 1j * rng.standard_normal((C_, n))
    eng.upload_vectors(X0)
    jac = (np.arange(C_) % 2).astype(np.uint8)
    psi = np.full(C_, 1e-19)
    t0 = time.perf_counter()
    X, st, it = eng.solve_shifted(np.zeros(C_, dtype=complex), psi, rng_key=np.arange(C_, dtype=np.uint64), method=_abi.METHOD_GMRES,
                                  use_jacobi=jac, RHS=None, rhs_shared=True)
    dt = time.perf_counter() - t0
    res = [float(np.linalg.norm(A @ X[c] - b) / np.linalg.norm(b)) for c in (0, 1, 2, 3)]
    emit(config="K4 dense Ax=b n=8192 c128 GMRES(20)x50 + Jacobi(half)", seconds=round(dt, 3), status_counts=np.bincount(st, minlength=4).tolist(),
         inner_iters_min=int(it.min()), inner_iters_max=int(it.max()), inner_iters_jacobi_mean=float(it[jac == 1].mean()),
         inner_iters_plain_mean=float(it[jac == 0].mean()), rel_residual_first4=res)

if "k5" in sections:
    # config 5: sparse CSC n=1M, ~20 nnz/row, GMRES / SpMV path (matrix replicated; candidates sharded across GPUs)
    n, C_ = 1_000_000, 8
    A = k5_sparse(n)
    eng.set_matrix(A)
    rng = np.random.default_rng(1)
    V = rng.random((C_, n)) + 1j * rng.random((C_, n)); V /= np.linalg.norm(V, axis=1, keepdims=True)
    eng.upload_vectors(V)
    import scipy.sparse as sp
    # (a) linear system A x = b (sigma = 0): the diagonally dominant operator GMRES(20) solves in about one cycle
    t0 = time.perf_counter()
    X, st, it = eng.solve_shifted(np.zeros(C_, dtype=complex), np.full(C_, 5e-19), rng_key=None, method=_abi.METHOD_GMRES, RHS=None)
    dt = time.perf_counter() - t0
    rel = float(np.linalg.norm(A @ X[0] - V[0]) / np.linalg.norm(V[0]))
    emit(config="K5 sparse n=1M nnz/row~21 c8 GMRES(20), linear system (sigma = 0)", seconds=round(dt, 4), status=st.tolist(),
         inner_iters=it.tolist(), rel_residual_c0=rel, ms_per_candidate_solve=round(dt / C_ * 1e3, 2))
    # (b) eigen step: shift = Rayleigh quotient of a random vector lies INSIDE the spectrum -> GMRES(20) x 50 stagnates, exactly
    # like scipy on the same operator (parity-tested at n = 3000); the reference's ladder then falls back
    lam, _ = eng.rq(C_=C_)
    t0 = time.perf_counter()
    X, st, it = eng.solve_shifted(lam, np.full(C_, 5e-19), rng_key=None, method=_abi.METHOD_GMRES, RHS=None)
    dt = time.perf_counter() - t0
    H = A - lam[0] * sp.eye(n, format="csc")
    rel = float(np.linalg.norm(H @ X[0] - V[0]) / np.linalg.norm(V[0]))
    emit(config="K5 sparse n=1M c8 GMRES(20)x50, eigen shift inside the spectrum", seconds=round(dt, 3), status=st.tolist(),
         inner_iters=it.tolist(), rel_residual_c0=rel, ms_per_inner_iteration_all_candidates=round(dt / max(1, int(it.max())) * 1e3, 3))
eng.close()

"""Seam-B host logic driven by the REAL reference objects (only where /root/reference exists, i.e. the build container):
``gpu_generation`` must run the reference's own ``MAUS_Solver`` / ``SolutionCandidate`` instances unmodified.  The numerics
come from the CPU FakeEngine (oracle arithmetic) because this tier has no GPU; what is checked is the attribute / method
contract and that one batched generation equals one generation of the reference's own loop."""
import os
import random
import sys
import warnings

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
from ref_loader import reference_available, load_reference, quiet, drive_generation   # noqa: E402

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")


def _build(ams, n, ncand, seed):
    np.random.seed(seed); random.seed(seed)
    M = ams.create_laplace_like_complex_eigen_for_MAUS(n, make_hermitian=False)
    return quiet(ams.MAUS_Solver, M, problem_type=ams.ProblemType.EIGENVALUE, initial_num_candidates=ncand,
                 global_convergence_tol=1e-9)


def test_first_generation_equals_reference_loop():
    from adaptive_matrix_solver_b200.population import gpu_generation
    from fake_engine import FakeEngine
    ams = load_reference(gmres_shim=True, name="ams_dropin_a")
    ref = _build(ams, 24, 10, 7)
    gpu = _build(ams, 24, 10, 7)
    ref_orig, gpu_orig = list(ref.candidates), list(gpu.candidates)      # _manage_candidates re-sorts / replaces the lists
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        quiet(drive_generation, ref, 1)
        quiet(gpu_generation, gpu, 1, FakeEngine())
    # _manage_candidates spawns from the global RNG, whose position differs (the reference burnt 2 N^2 draws per step on
    # the Psi perturbation): compare the candidates that existed before the generation
    for a, b in zip(ref_orig, gpu_orig):
        assert a.state == b.state and a.stuck_counter == b.stuck_counter
        assert abs(a.lambda_k - b.lambda_k) <= 1e-10 * abs(a.lambda_k) + 1e-12
        assert abs(a.residual_k - b.residual_k) <= 1e-9 * a.residual_k + 1e-12
        ph = np.vdot(b.v_k, a.v_k); ph /= abs(ph)
        assert np.abs(b.v_k * ph - a.v_k).max() <= 1e-9
        assert complex(a.alpha_local_step) == complex(b.alpha_local_step)
        assert len(a.residual_history) == len(b.residual_history) and len(a.param_history) == len(b.param_history)


def test_generations_converge_to_true_eigenvalues_with_reference_objects():
    from adaptive_matrix_solver_b200.population import gpu_generation
    from fake_engine import FakeEngine
    ams = load_reference(gmres_shim=True, name="ams_dropin_b")
    s = _build(ams, 16, 12, 3)
    eng = FakeEngine()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for it in range(1, 26):
            quiet(gpu_generation, s, it, eng)
    assert s.num_distinct_converged_solutions >= 1           # the reference's own diagnostics ran on our write-back
    ev = np.linalg.eigvals(s.M)
    for lam, v in s.converged_solutions:
        assert np.abs(ev - lam).min() < 1e-7
        assert np.linalg.norm(s.M @ v - lam * v) < 1e-7


def test_seam_a_name_rebinding():
    """install_dropin rebinds the module-global name the reference resolves on every step (AMS:224)."""
    import adaptive_matrix_solver_b200 as pkg
    ams = load_reference(gmres_shim=True, name="ams_dropin_c")
    orig = ams.InverseIterateSolver
    pkg.install_dropin(ams, engine=object())
    assert ams.InverseIterateSolver is pkg.GpuInverseIterateSolver and orig is not pkg.GpuInverseIterateSolver
    s = ams.InverseIterateSolver(5, 1e-20, 25, 'iterative_gmres', True)
    assert (s.N, s.max_attempts, s.preferred_method, s.fallback_method, s.is_sparse) == (5, 25, 'iterative_gmres', 'direct_solve', True)
    pkg.GpuInverseIterateSolver.bind_engine(None)


def test_device_vdot_leaves_the_reference_dedup_decisions_unchanged():
    """the reference's own _update_global_diagnostics / _manage_candidates, run with np.vdot answered from
    the Gram matrix, must take exactly the decisions they take with numpy's vdot."""
    from adaptive_matrix_solver_b200.dedup import device_vdot
    from fake_engine import FakeEngine
    ams = load_reference(gmres_shim=True, name="ams_dropin_c")
    n = 12

    def population(seed):
        s = _build(ams, n, 14, seed)
        w, Vr = np.linalg.eig(np.asarray(s.M))
        rng = np.random.default_rng(5)
        for k, c in enumerate(s.candidates):
            j = k % 5                                         # 5 distinct eigenpairs, several candidates on each
            ph = np.exp(1j * rng.uniform(0, 6.28))
            c.lambda_k = w[j]; c.v_k = (Vr[:, j] / np.linalg.norm(Vr[:, j])) * ph
            c.residual_k = 1e-12 * (1 + k); c.w_k = 1.0
            c.state = ams.SolutionCandidate.State.CONVERGED if k < 11 else ams.SolutionCandidate.State.EXPLORING
        return s

    ref, dev = population(3), population(3)
    quiet(ref._update_global_diagnostics, 1)
    np.random.seed(9); random.seed(9)
    quiet(ref._manage_candidates, 1)
    with device_vdot(ams, dev.candidates, FakeEngine()) as px:
        quiet(dev._update_global_diagnostics, 1)
    assert px.hits > 0
    np.random.seed(9); random.seed(9)
    with device_vdot(ams, dev.candidates, FakeEngine()) as px2:
        quiet(dev._manage_candidates, 1)
    assert px2.hits > 0
    assert ams.np is np                                              # the proxy is gone
    assert ref.num_distinct_converged_solutions == dev.num_distinct_converged_solutions == 5
    assert [c.state for c in ref.candidates] == [c.state for c in dev.candidates]
    assert len(ref.candidates) == len(dev.candidates)
    assert np.allclose([c.lambda_k for c in ref.candidates if c.lambda_k is not None][:5],
                       [c.lambda_k for c in dev.candidates if c.lambda_k is not None][:5])


def test_hermitian_shortcut_with_one_shared_eigh_equals_the_reference_loop():
    """the dense Hermitian shortcut (AMS:155-186) runs eigh ONCE for the group and matches every candidate on
    the device; the outcome per candidate must equal the reference's own per-candidate eigh."""
    from adaptive_matrix_solver_b200.population import step_population
    from fake_engine import FakeEngine
    ams = load_reference(gmres_shim=True, name="ams_dropin_h")

    def build(seed):
        np.random.seed(seed); random.seed(seed)
        M = ams.create_laplace_like_complex_eigen_for_MAUS(18, make_hermitian=True)
        return quiet(ams.MAUS_Solver, M, problem_type=ams.ProblemType.EIGENVALUE, initial_num_candidates=9, global_convergence_tol=1e-9)

    ref, dev = build(11), build(11)
    assert ref.problem_knowledge.get('is_hermitian', False)
    for c in ref.candidates:
        quiet(c.update_solution_step, ref.M, ref.b, ref.strat_params, ref.problem_knowledge)
    eng = FakeEngine()
    calls = {"n": 0}
    import scipy.linalg as sla
    real_eigh = sla.eigh

    def counting_eigh(*a, **k):
        calls["n"] += 1
        return real_eigh(*a, **k)
    sla.eigh = counting_eigh
    try:
        n_stepped = quiet(step_population, dev.candidates, dev.M, dev.b, dev.strat_params, dev.problem_knowledge, eng)
    finally:
        sla.eigh = real_eigh
    assert n_stepped == 9 and calls["n"] == 1                        # one factorization for the whole population
    for a, b_ in zip(ref.candidates, dev.candidates):
        assert a.state == b_.state == ams.SolutionCandidate.State.CONVERGED
        assert a.lambda_k == b_.lambda_k and type(a.lambda_k) is type(b_.lambda_k)
        assert np.array_equal(a.v_k, b_.v_k)
        assert abs(a.residual_k - b_.residual_k) <= 1e-13
        assert a.w_k == b_.w_k == 1.0 and a.stuck_counter == b_.stuck_counter == 0
        assert len(a.residual_history) == len(b_.residual_history) and len(a.param_history) == len(b_.param_history)


def test_generation_with_device_dedup_equals_generation_without():
    """dedup.gpu_generation_dedup is gpu_generation with the similarity tests answered from the Gram matrix: every
    generation of a run must leave the same population behind."""
    from adaptive_matrix_solver_b200.population import gpu_generation
    from adaptive_matrix_solver_b200.dedup import gpu_generation_dedup
    from fake_engine import FakeEngine
    ams = load_reference(gmres_shim=True, name="ams_dropin_d")
    a, b_ = _build(ams, 16, 12, 3), _build(ams, 16, 12, 3)
    ea, eb = FakeEngine(), FakeEngine()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for it in range(1, 29):
            np.random.seed(100 + it); random.seed(100 + it)
            quiet(gpu_generation, a, it, ea)
            np.random.seed(100 + it); random.seed(100 + it)
            quiet(gpu_generation_dedup, ams, b_, it, eb)
            assert len(a.candidates) == len(b_.candidates)
            assert [c.state for c in a.candidates] == [c.state for c in b_.candidates]
            assert a.num_distinct_converged_solutions == b_.num_distinct_converged_solutions
    assert a.num_distinct_converged_solutions >= 1
    for ca, cb in zip(a.candidates, b_.candidates):
        assert ca.lambda_k == cb.lambda_k and np.array_equal(ca.v_k, cb.v_k)


def test_initial_strategy_mirrors_the_reference_for_its_own_scenarios():
    """diagnostics.initial_strategy (AMS:405-421 as a pure function) against the strategy the REAL MAUS_Solver.__init__ derives,
    for matrices on each side of the cond thresholds and every problem type."""
    from adaptive_matrix_solver_b200.diagnostics import initial_strategy
    ams = load_reference(gmres_shim=True, name="ams_diag")
    rng = np.random.default_rng(0)
    G = rng.standard_normal((40, 40)) + 1j * rng.standard_normal((40, 40))
    U, _, Vh = np.linalg.svd(G)
    mats = [G, (U * np.logspace(0, -8, 40)) @ Vh, (U * np.logspace(0, -13, 40)) @ Vh, (U * np.r_[np.ones(39), 0.0]) @ Vh]
    for M in mats:
        for pt in (ams.ProblemType.EIGENVALUE, ams.ProblemType.SOLVE_LINEAR_SYSTEM, ams.ProblemType.SVD):
            np.random.seed(1); random.seed(1)
            b = np.ones(40, dtype=np.complex128) if pt == ams.ProblemType.SOLVE_LINEAR_SYSTEM else None
            s = quiet(ams.MAUS_Solver, M, problem_type=pt, b_vector=b, initial_num_candidates=2, global_convergence_tol=1e-8)
            strat, know = initial_strategy(s.diag_info, pt.name, 1e-8)
            for k, v in strat.items():
                assert s.strat_params[k] == v, (pt, k)
            for k, v in know.items():
                assert s.problem_knowledge[k] == v, (pt, k)


def test_install_diagnostics_drives_the_reference_constructor():
    """diagnostics.install_diagnostics rebinds MAUS_Solver._diagnose_matrix_initial (AMS:345 calls it through self): the REAL
    constructor must build the same diag_info / strategy / problem_knowledge from the engine-backed diagnosis as from its own
    numpy one (dense, Hermitian, ill-conditioned and sparse inputs)."""
    import scipy.sparse as sp
    from adaptive_matrix_solver_b200.diagnostics import install_diagnostics
    from fake_engine import FakeEngine
    rng = np.random.default_rng(3)
    G = rng.standard_normal((30, 30)) + 1j * rng.standard_normal((30, 30))
    U, _, Vh = np.linalg.svd(G)
    mats = [G, G + G.conj().T, (U * np.logspace(0, -9, 30)) @ Vh, sp.random(30, 30, density=0.2, random_state=1, format="csc") + sp.eye(30, format="csc")]
    for M in mats:
        ref_mod = load_reference(gmres_shim=True, name="ams_diag_ref")
        gpu_mod = install_diagnostics(load_reference(gmres_shim=True, name="ams_diag_gpu"), FakeEngine())
        np.random.seed(2); random.seed(2)
        a = quiet(ref_mod.MAUS_Solver, M, problem_type=ref_mod.ProblemType.EIGENVALUE, initial_num_candidates=2)
        np.random.seed(2); random.seed(2)
        b = quiet(gpu_mod.MAUS_Solver, M, problem_type=gpu_mod.ProblemType.EIGENVALUE, initial_num_candidates=2)
        for k in ("is_hermitian", "is_complex_symmetric", "is_sparse_init", "is_singular"):
            assert a.diag_info[k] == b.diag_info[k], k
        ca, cb = a.diag_info["condition_number"], b.diag_info["condition_number"]
        assert (np.isinf(ca) and np.isinf(cb)) or abs(ca - cb) <= 1e-8 * ca
        assert a.strat_params == b.strat_params and a.problem_knowledge == b.problem_knowledge

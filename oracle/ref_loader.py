"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (never shipped, never on the GPU box).

Imports /root/reference/Adaptive_Matrix_Solver_0.1.py (AMS) without editing it, for two purposes only:
  * oracle/gen_golden.py  -- generate the golden traces committed under tests/golden/
  * tests (not gpu)       -- cross-check oracle/maus_oracle.py against the reference when it is present

Work-arounds applied from the OUTSIDE:
  * AMS:89 calls ``spla.gmres(..., tol=1e-8)``; scipy >= 1.14 renamed the keyword to ``rtol`` so the call
    raises TypeError, is swallowed at AMS:98 and the GMRES branch silently falls back to the direct solver.
    ``load_reference(gmres_shim=True)`` replaces the module attribute ``ams.spla`` by a proxy forwarding
    ``tol`` as ``rtol``.  ``gmres_shim=False`` keeps the as-shipped behaviour.
  * AMS:583 reads the undefined global ``target_sols_final``; ``evolve`` is therefore never called here --
    ``drive_generation`` performs the four calls of AMS:573-577 itself.
"""
import importlib.util
import io
import os
import contextlib

REFERENCE_FILE = "/root/reference/Adaptive_Matrix_Solver_0.1.py"


def reference_available():
    return os.path.isfile(REFERENCE_FILE)


class _SplaProxy:
    """scipy.sparse.linalg with gmres(tol=...) mapped onto gmres(rtol=...)."""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    def gmres(self, A, b, x0=None, tol=None, **kw):
        if tol is not None:
            kw["rtol"] = tol
        return self._real.gmres(A, b, x0=x0, **kw)


def load_reference(gmres_shim=True, name="ams_reference"):
    if not reference_available():
        raise FileNotFoundError(REFERENCE_FILE)
    spec = importlib.util.spec_from_file_location(name, REFERENCE_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # __name__ != '__main__', the scenario block does not run
    if gmres_shim:
        mod.spla = _SplaProxy(mod.spla)
    return mod


def quiet(fn, *a, **kw):
    """Run fn with the reference's prints swallowed."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def drive_generation(solver, iteration, on_step=None):
    """One generation exactly as the body of the loop at AMS:572-577 (evolve itself crashes, AMS:583).

    ``on_step(candidate, before_snapshot)`` is called around every update_solution_step when given.
    """
    ams_State = type(solver.candidates[0]).State if solver.candidates else None
    solver._update_global_diagnostics(iteration)
    solver._adjust_global_strategy(iteration)
    for cand in solver.candidates:
        if cand.state not in (ams_State.CONVERGED, ams_State.RETIRED):
            if on_step is not None:
                on_step(cand, solver)
            else:
                cand.update_solution_step(solver.M, solver.b, solver.strat_params, solver.problem_knowledge)
    solver._manage_candidates(iteration)

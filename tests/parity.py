"""Comparison helpers shared by the GPU parity tests.  Tolerances (BASELINE.json north_star): 1e-10 relative in
complex128 for eigenvalues, vectors (up to phase) and residuals, with the absolute floor the domain needs near
convergence -- lambda and the stale-lambda residual (AMS:297) are differences of O(||A||) quantities, so rounding
in either implementation moves them by O(eps*||A||) regardless of their size."""
import numpy as np

RTOL = 1e-10


def anorm(A):
    import scipy.sparse as sp
    if sp.issparse(A):
        return float(abs(A).sum(axis=1).max())
    return float(np.abs(A).sum(axis=1).max())


def phase_align(v, ref):
    ph = np.vdot(v, ref)
    if abs(ph) == 0:
        return v
    return v * (ph / abs(ph))


def assert_scalar_close(a, b, floor, what=""):
    a = complex(a); b = complex(b)
    if np.isnan(a.real) or np.isnan(b.real):
        assert np.isnan(a.real) == np.isnan(b.real), what
        return
    assert abs(a - b) <= RTOL * max(abs(a), abs(b)) + floor, f"{what}: {a} vs {b} (floor {floor:g})"


def vec_err_up_to_phase(v, ref):
    return float(np.abs(phase_align(v, ref) - ref).max() / max(np.abs(ref).max(), 1e-300))

"""Device-side similarity tests for the reference's converged-solution dedup and survivor pruning.

``MAUS_Solver._update_global_diagnostics`` (AMS:427-453) and ``_manage_candidates`` (AMS:507-524) decide whether two
CONVERGED candidates describe the same solution with ``abs(np.vdot(v_i, v_j)) > 0.999`` inside nested Python loops:
O(conv^2) host dot products of length N per generation (conv = 256, N = 4096: ~0.5 s, as long as the whole GPU generation).
Here all of them are ONE device pass (``maus_gram``: G = V^H V) and the reference's code runs UNCHANGED: for the duration of
the two calls the module-global ``np`` of the reference is a proxy whose ``vdot`` answers from G when both arguments are
vectors of converged candidates, and falls through to numpy for anything else (the same outside-the-file technique as the
``tol -> rtol`` shim on ``ams.spla``).
"""
import contextlib

import numpy as np

_CONVERGED, _RETIRED = "CONVERGED", "RETIRED"


class _NumpyWithGram:
    """numpy look-alike for the reference module: everything forwards to numpy except ``vdot`` on registered vectors."""

    def __init__(self, real_np, table, grams):
        self._np, self._table, self._grams = real_np, table, grams
        self.hits = 0
        self.misses = 0

    def __getattr__(self, name):
        return getattr(self._np, name)

    def vdot(self, a, b):
        ka, kb = self._table.get(id(a)), self._table.get(id(b))
        if ka is not None and kb is not None and ka[0] == kb[0] and ka[2] is a and kb[2] is b:
            self.hits += 1
            return self._grams[ka[0]][ka[1], kb[1]]
        self.misses += 1
        return self._np.vdot(a, b)


def similarity_tables(candidates, engine):
    """Gram matrices of the vectors the dedup tests compare: EIGENVALUE -> v_k; SVD -> u_k and right_v_k (two groups).
    Returns (table {id(array): (group, row, array)}, grams {group: G})."""
    groups = {}
    for c in candidates:
        if c.state.name != _CONVERGED:
            continue
        pt = c.problem_type.name
        if pt == "EIGENVALUE":
            fields = (("v", c.v_k),)
        elif pt == "SVD":
            fields = (("u", c.u_k), ("rv", c.right_v_k))
        else:
            continue                     # SOLVE_LINEAR_SYSTEM compares ||x_i - x_0|| (AMS:439-441): not a dot-product test
        for g, vec in fields:
            if isinstance(vec, np.ndarray) and vec.ndim == 1 and vec.size:
                groups.setdefault((g, vec.size), []).append(vec)
    table, grams = {}, {}
    for key, vecs in groups.items():
        if len(vecs) < 2:
            continue
        V = np.ascontiguousarray(np.stack(vecs), dtype=np.complex128)
        grams[key] = engine.gram(V)
        for r, vec in enumerate(vecs):
            table[id(vec)] = (key, r, vec)          # holding `vec` keeps the id unique while the table lives
    return table, grams


@contextlib.contextmanager
def device_vdot(ams_module, candidates, engine):
    """Within the block, ``np.vdot`` calls made BY THE REFERENCE MODULE on converged candidates' vectors are answered from
    the device Gram matrix.  Yields the proxy (``.hits`` / ``.misses`` count the answered / forwarded calls)."""
    table, grams = similarity_tables(candidates, engine)
    real = ams_module.np
    proxy = _NumpyWithGram(real, table, grams)
    ams_module.np = proxy
    try:
        yield proxy
    finally:
        ams_module.np = real


def gpu_generation_dedup(ams_module, maus_solver, iteration, engine):
    """``population.gpu_generation`` with the similarity tests of both host phases on the device (one Gram pass each: the
    set of converged candidates changes in between)."""
    from .population import step_population
    with device_vdot(ams_module, maus_solver.candidates, engine):
        maus_solver._update_global_diagnostics(iteration)
    maus_solver._adjust_global_strategy(iteration)
    n = step_population(maus_solver.candidates, maus_solver.M, maus_solver.b, maus_solver.strat_params,
                        maus_solver.problem_knowledge, engine)
    with device_vdot(ams_module, maus_solver.candidates, engine):
        maus_solver._manage_candidates(iteration)
    return n

"""Seam A: ``GpuInverseIterateSolver`` -- signature-identical replacement of the reference's
``InverseIterateSolver`` (AMS:30-104).  The reference resolves that name through its module globals on every
step (AMS:224), so ``ams.InverseIterateSolver = GpuInverseIterateSolver`` swaps the solver with zero edits.

The retry / fallback ladder (AMS:43, 98-104) stays on the host, line for line; every attempt is ONE call into
the CUDA library (``maus_solve_shifted``, C = 1).  Errors follow the reference: ``RuntimeError`` when all attempts
fail (AMS:104).  ``A_target`` / ``b_rhs`` are never mutated; the result is a fresh array.
"""
import itertools

import numpy as np

from . import _abi
from .constants import LU_MAX_N, psi_magnitude
from .engine import MausEngine

_key_counter = itertools.count(1)


def _is_sparse(A):
    try:
        import scipy.sparse as sp
        return sp.issparse(A)
    except Exception:  # pragma: no cover
        return False


class GpuInverseIterateSolver:
    _shared_engine = None

    @classmethod
    def bind_engine(cls, engine):
        cls._shared_engine = engine

    def __init__(self, N, base_psi_epsilon, max_attempts, preferred_method='direct_solve', is_sparse=False):
        self.N = N
        self.base_psi_epsilon = base_psi_epsilon
        self.max_attempts = max_attempts
        self.preferred_method = preferred_method
        self.fallback_method = 'iterative_gmres' if preferred_method == 'direct_solve' else 'direct_solve'   # AMS:36
        self.is_sparse = is_sparse

    def _engine(self):
        if GpuInverseIterateSolver._shared_engine is None:
            GpuInverseIterateSolver._shared_engine = MausEngine(0)
        return GpuInverseIterateSolver._shared_engine

    def solve(self, A_target, b_rhs, candidate_stuck_counter):
        eng = self._engine()
        sparse_in = _is_sparse(A_target)
        uploaded = None            # 'sparse' | 'dense'
        rhs = np.ascontiguousarray(b_rhs, dtype=np.complex128)
        num_psi_attempts = 0
        method = self.preferred_method
        while num_psi_attempts < self.max_attempts:                                             # AMS:43
            psi = psi_magnitude(self.base_psi_epsilon, num_psi_attempts, candidate_stuck_counter)
            status = None
            if method == 'direct_solve':
                if A_target.shape[0] > LU_MAX_N:
                    status = _abi.ST_ZERO_PIVOT      # beyond the batched LU (sparse or dense): behaves like a failed try
                else:
                    if uploaded != 'dense':
                        eng.set_matrix(A_target.toarray() if sparse_in else A_target)
                        uploaded = 'dense'
                    # dense: random Psi perturbation (AMS:49); sparse input: psi*I only (AMS:47)
                    key = None if self.is_sparse else [(next(_key_counter) << 8) | (num_psi_attempts & 0xff)]
                    X, st, _ = eng.solve_shifted([0j], [complex(psi).real], rng_key=key, method=_abi.METHOD_LU,
                                                 RHS=rhs[None, :])
                    status = int(st[0])
            elif method == 'iterative_gmres':
                want = 'sparse' if sparse_in else 'dense'
                if uploaded != want:
                    eng.set_matrix(A_target)
                    uploaded = want
                key = None if self.is_sparse else [(next(_key_counter) << 8) | (num_psi_attempts & 0xff)]
                X, st, _ = eng.solve_shifted([0j], [complex(psi).real], rng_key=key, method=_abi.METHOD_GMRES,
                                             use_jacobi=[1 if candidate_stuck_counter > 1 else 0],       # AMS:65
                                             RHS=rhs[None, :])
                status = int(st[0])
            else:
                status = _abi.ST_NONFINITE      # AMS:92: the ValueError is swallowed by the ladder's own except (AMS:98)
            if status == _abi.ST_OK:
                return X[0].copy(), num_psi_attempts                                             # AMS:97
            # AMS:98-103
            if method == self.preferred_method and self.preferred_method != self.fallback_method and num_psi_attempts == 0:
                method = self.fallback_method
                num_psi_attempts = 0
                continue
            num_psi_attempts += 1
        raise RuntimeError(f"InverseIterateSolver failed all {self.max_attempts} attempts for "
                           f"{self.preferred_method} and {self.fallback_method}.")              # AMS:104

"""Heuristic constants of the reference that feed the hot path (AMS:16-26); mirrored exactly, including dtypes."""
import numpy as np

PSI_EPSILON_BASE = np.complex128(1e-20)       # GLOBAL_DEFAULT_PSI_EPSILON_BASE, AMS:16
ALPHA_V_INITIAL = np.complex128(0.01)         # GLOBAL_DEFAULT_ALPHA_V_INITIAL, AMS:17
MAX_PSI_ATTEMPTS = 25                         # GLOBAL_MAX_PSI_ATTEMPTS, AMS:18
MAX_STUCK_FOR_RETIREMENT = 8                  # GLOBAL_MAX_STUCK_FOR_RETIREMENT, AMS:19
SIGMA_SIMILARITY_TOL_ABS = 1e-6               # GLOBAL_SIGMA_SIMILARITY_TOL_ABS, AMS:23
CONVERGENCE_RESIDUAL_TOL = 1e-8               # GLOBAL_CONVERGENCE_RESIDUAL_TOL, AMS:25
LU_MAX_N = 8192                               # largest order the batched LU accepts (csrc/lu.cuh)


def psi_magnitude(base_psi_epsilon, num_psi_attempts, candidate_stuck_counter):
    """AMS:44 -- complex128 scalar with zero imaginary part."""
    return base_psi_epsilon * (10 ** (num_psi_attempts / 2.0)) * (10 ** (candidate_stuck_counter / 3.0))

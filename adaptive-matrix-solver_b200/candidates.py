"""Stand-alone candidate objects with the reference's public field layout (AMS:113-126), for users (and bench.py)
that drive ``step_population`` without importing the reference.  The reference's own ``SolutionCandidate`` objects
work with ``step_population`` as they are; this class only mirrors their attribute names and the two methods the step
calls (``initialize_random_solution`` AMS:129-143, ``get_current_solution_params`` AMS:333-337)."""
import random
from enum import Enum

import numpy as np

from .constants import ALPHA_V_INITIAL


class ProblemType(Enum):                 # AMS:10-13
    EIGENVALUE = 1
    SOLVE_LINEAR_SYSTEM = 2
    SVD = 3


class Candidate:
    class State(Enum):                   # AMS:109-110
        EXPLORING = 1
        REFINING = 2
        STUCK = 3
        CONVERGED = 4
        RETIRED = 5

    _candidate_id_counter = 0

    def __init__(self, problem_matrix, problem_type, N_diag, initial_lambda=None, initial_v=None, initial_x=None,
                 initial_u=None, initial_sigma=None, initial_weight=0.01):
        self.id = Candidate._candidate_id_counter
        Candidate._candidate_id_counter += 1
        self.N_diag = N_diag
        self.M_rows, self.M_cols = problem_matrix.shape
        self.problem_type = problem_type
        self.problem_matrix = problem_matrix
        self.b_vector = None
        self.lambda_k = initial_lambda
        self.v_k = initial_v
        self.x_k = initial_x
        self.sigma_k = initial_sigma
        self.u_k = initial_u
        self.right_v_k = initial_v
        self.state = Candidate.State.EXPLORING
        self.w_k = initial_weight
        self.residual_k = float('inf')
        self.prev_residual = float('inf')
        self.alpha_local_step = ALPHA_V_INITIAL
        self.stuck_counter = 0
        self.local_psi_retries_needed = 0
        self.num_resets = 0
        self.param_history = []
        self.residual_history = []
        if (problem_type == ProblemType.EIGENVALUE and initial_v is None) or \
                (problem_type == ProblemType.SOLVE_LINEAR_SYSTEM and initial_x is None) or \
                (problem_type == ProblemType.SVD and (initial_u is None or initial_v is None)):
            self.initialize_random_solution()

    def initialize_random_solution(self):
        def unit_random(n=self.N_diag):
            v = (np.random.rand(n) + 1j * np.random.rand(n)).astype(np.complex128)
            nv = np.linalg.norm(v)
            return v / nv if nv > 1e-10 else np.full(n, 1.0 / np.sqrt(n), dtype=np.complex128)

        if self.problem_type == ProblemType.EIGENVALUE:
            self.v_k = unit_random()
            self.lambda_k = (random.random() * 5 - 2.5 + 1j * (random.random() * 5 - 2.5))
        elif self.problem_type == ProblemType.SOLVE_LINEAR_SYSTEM:
            self.x_k = unit_random() * random.uniform(0.1, 10.0)
        elif self.problem_type == ProblemType.SVD:
            self.u_k = unit_random(self.M_rows)
            self.right_v_k = unit_random(self.M_cols)
            self.sigma_k = 1.0
        self.param_history.append(self.get_current_solution_params())
        self.residual_history.append(self.residual_k)

    def get_current_solution_params(self):
        if self.problem_type == ProblemType.EIGENVALUE:
            return (self.lambda_k, self.v_k)
        if self.problem_type == ProblemType.SOLVE_LINEAR_SYSTEM:
            return (self.x_k,)
        if self.problem_type == ProblemType.SVD:
            return (self.sigma_k, self.u_k, self.right_v_k)
        return None

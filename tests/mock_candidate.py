"""A stand-in for the reference's ``SolutionCandidate`` (AMS:107-143, 333-337) with exactly the attributes and methods
``step_population`` touches, so the Seam-B host logic can be tested where /root/reference does not exist (GPU box).
Random initialisation follows the oracle's restatement of AMS:129-143."""
from enum import Enum

import numpy as np

from oracle import maus_oracle as mo


class ProblemType(Enum):
    EIGENVALUE = 1
    SOLVE_LINEAR_SYSTEM = 2
    SVD = 3


class MockCandidate:
    class State(Enum):
        EXPLORING = 1; REFINING = 2; STUCK = 3; CONVERGED = 4; RETIRED = 5

    _next_id = 0

    def __init__(self, problem_matrix, problem_type, N):
        self.id = MockCandidate._next_id
        MockCandidate._next_id += 1
        self.N_diag = N
        self.M_rows, self.M_cols = problem_matrix.shape
        self.problem_type = problem_type
        self.problem_matrix = problem_matrix
        self.b_vector = None
        self.lambda_k = None; self.v_k = None; self.x_k = None
        self.u_k = None; self.right_v_k = None; self.sigma_k = None
        self.state = MockCandidate.State.EXPLORING
        self.w_k = 0.01
        self.residual_k = float('inf'); self.prev_residual = float('inf')
        self.alpha_local_step = mo.ALPHA_V_INITIAL
        self.stuck_counter = 0; self.local_psi_retries_needed = 0; self.num_resets = 0
        self.param_history = []; self.residual_history = []
        self.initialize_random_solution()

    def initialize_random_solution(self):
        c = mo.CandState(problem_type=self.problem_type.value, N=self.N_diag, M_rows=self.M_rows, M_cols=self.M_cols)
        mo.initialize_random_solution(c)
        self.v_k, self.lambda_k, self.x_k = c.v_k, c.lambda_k, c.x_k
        self.u_k, self.right_v_k, self.sigma_k = c.u_k, c.right_v_k, c.sigma_k
        self.param_history.append(self.get_current_solution_params())
        self.residual_history.append(self.residual_k)

    def get_current_solution_params(self):
        if self.problem_type == ProblemType.EIGENVALUE:
            return (self.lambda_k, self.v_k)
        if self.problem_type == ProblemType.SVD:
            return (self.sigma_k, self.u_k, self.right_v_k)
        return (self.x_k,)

    # ---- conversion helpers for the tests -------------------------------------------------------------------
    def load(self, s):
        """s: dict from golden_io.Golden.side(...)"""
        self.lambda_k = s["lam"]; self.v_k = None if s["v"] is None else s["v"].copy()
        self.x_k = None if s["x"] is None else s["x"].copy()
        self.state = MockCandidate.State(s["state"]); self.w_k = s["w"]
        self.residual_k = s["res"]; self.prev_residual = s["prev"]; self.alpha_local_step = s["alpha"]
        self.stuck_counter = s["stuck"]; self.local_psi_retries_needed = s["retries"]; self.num_resets = s["resets"]
        self.param_history = [None] * s["hist"]; self.residual_history = [None] * s["hist"]
        if "u" in s:
            self.u_k = s["u"].copy(); self.right_v_k = s["rv"].copy(); self.sigma_k = s["sigma"]
        return self

    def to_oracle(self):
        c = mo.CandState(problem_type=self.problem_type.value, N=self.N_diag, M_rows=self.M_rows, M_cols=self.M_cols)
        c.u_k = None if self.u_k is None else self.u_k.copy()
        c.right_v_k = None if self.right_v_k is None else self.right_v_k.copy(); c.sigma_k = self.sigma_k
        c.lambda_k = self.lambda_k; c.v_k = None if self.v_k is None else self.v_k.copy()
        c.x_k = None if self.x_k is None else self.x_k.copy()
        c.state = self.state.value; c.w_k = self.w_k; c.residual_k = self.residual_k
        c.prev_residual = self.prev_residual; c.alpha_local_step = self.alpha_local_step
        c.stuck_counter = self.stuck_counter; c.local_psi_retries_needed = self.local_psi_retries_needed
        c.num_resets = self.num_resets; c.history_len = len(self.residual_history)
        return c

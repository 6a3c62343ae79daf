"""CPU stand-in for MausEngine used ONLY by the world_size-2 gloo tests of the host-side sharding logic (there is no
GPU in the CPU test tier).  Numerics come from the oracle; it mimics the subset of the engine API step_population uses."""
import numpy as np
import scipy.linalg as sla

from oracle import maus_oracle as mo


class FakeEngine:
    def __init__(self):
        self.n = 0
        self.is_sparse = False
        self.generation = 0
        self.A = [None, None]
        self.b = None
        self.calls = 0
        self.fail_ids = set()        # candidate ids whose every solve attempt fails (exercises the ladder + re-initialisation)
        self.matrix_epoch = [0, 0]

    def set_matrix(self, A, slot=0):
        self.A[slot] = np.asarray(A, dtype=np.complex128)
        if slot == 0:
            self.n = self.A[0].shape[0]
            self.A[1] = None
            self.matrix_epoch[1] += 1
        self.matrix_epoch[slot] += 1

    def upload_vectors(self, V):
        pass

    # set-up diagnostics (numpy stand-ins of maus_diag_dense / maus_cond2_estimate)
    def diag_dense(self, rtol=1e-5, atol=1e-8):
        A = self.A[0]
        return int(np.count_nonzero(A)), bool(np.allclose(A, A.conj().T, rtol=rtol, atol=atol)), bool(np.allclose(A, A.T, rtol=rtol, atol=atol))

    def cond2_estimate(self, power_iters=40, inverse_iters=8, start=None):
        s = np.linalg.svd(self.A[0], compute_uv=False)
        return float(s[0]), float(s[-1]), 0 if s[-1] > 0 else 1

    def solve_shifted(self, sigma, psi, rng_key=None, method=0, use_jacobi=None, RHS=None, rhs_shared=False, want_x=True):
        C_ = len(sigma)          # only reached for candidates in fail_ids: every attempt of the ladder fails
        return None, np.full(C_, 1, dtype=np.int32), np.zeros(C_, dtype=np.int32)

    def project(self, Ec, V):
        return (np.asarray(Ec).T @ np.asarray(V, dtype=np.complex128).T).T

    def residual(self, problem_type, V=None, lam=None, C_=None, res_slot=0):
        A = self.A[0]
        return np.array([np.linalg.norm(A @ V[c] - lam[c] * V[c]) for c in range(V.shape[0])])

    def gram(self, V):
        V = np.asarray(V, dtype=np.complex128)
        return V.conj() @ V.T

    def set_rhs(self, b):
        self.b = np.asarray(b, dtype=np.complex128)

    def step(self, problem_type, alpha, psi, V=None, rng_key=None, method=0, use_jacobi=None, res_slot=0, out=None):
        C_ = len(alpha)
        A = self.A[0]
        Ares = self.A[res_slot] if self.A[res_slot] is not None else A
        lam = np.zeros(C_, dtype=np.complex128); resid = np.zeros(C_); mixn = np.zeros(C_)
        status = np.zeros(C_, dtype=np.int32); iters = np.zeros(C_, dtype=np.int32)
        for c in range(C_):
            if rng_key is not None and (int(rng_key[c]) >> 32) in self.fail_ids:
                if problem_type == 1:
                    lam[c] = mo.rayleigh_quotient(A, V[c])
                status[c] = 1
                continue
            if problem_type == 1:
                lam[c] = mo.rayleigh_quotient(A, V[c])
                x = mo.shifted_solve_dense(A, lam[c], psi[c], V[c])
                v2, nv = mo.mix_normalise(V[c], x, alpha[c])
                V[c] = v2; mixn[c] = nv
                resid[c] = mo.residual_eigen(Ares, V[c], lam[c])
            else:
                x = sla.solve(A + psi[c] * np.eye(self.n), self.b)
                V[c] = (1.0 - alpha[c]) * V[c] + alpha[c] * x
                resid[c] = mo.residual_linear(Ares, V[c], self.b)
        self.generation += 1
        self.calls += 1
        return dict(lam=lam, resid=resid, mixnorm=mixn, status=status, iters=iters)

"""Helpers to read tests/golden/*.npz (written by oracle/gen_golden.py from the real reference)."""
import json
import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["eig8", "eig100", "lin5_shim", "lin5_shipped", "gmres64", "speig200", "fail6", "svd5x4", "svd40x28"]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.z = z
        self.meta = json.loads(str(z["meta"]))
        self.n_steps = len(z["seed"])
        if self.meta["A_format"] == "csc":
            n = len(z["A_indptr"]) - 1
            self.A = sp.csc_matrix((z["A_data"], z["A_indices"], z["A_indptr"]), shape=(n, n))
        else:
            self.A = z["A"]
        self.n = self.A.shape[0]
        self.b = z["b"] if "b" in z.files else None
        self.A_ctor = z["A_ctor"] if "A_ctor" in z.files else None
        self.problem_type = self.meta["problem_type"]
        self.gmres_mode = self.meta["gmres_mode"]

    def ctor_matrix(self, i):
        if self.A_ctor is None or self.z["ctor_is_current"][i]:
            return self.A
        return self.A_ctor

    def strat(self, i):
        return dict(self.meta["steps"][i]["strat"])

    def know(self, i):
        st = self.meta["steps"][i]
        return dict(local_solver_preference=st["pref"], is_sparse_problem=st["sparse"], is_hermitian=False)

    def side(self, side, i):
        z = self.z
        g = lambda k: z[f"{side}_{k}"][i]
        alpha = g("alpha")
        alpha = np.complex128(alpha) if g("alpha_is_complex") else float(alpha.real)
        v = g("v"); x = g("x")
        extra = {}
        if self.problem_type == 3:
            extra = dict(u=g("u")[:self.A.shape[0]].copy(), rv=g("rv")[:self.A.shape[1]].copy(), sigma=float(g("sigma")))
        return dict(**extra, lam=complex(g("lam")), v=None if np.isnan(v).all() and self.problem_type != 1 else v,
                    x=None if np.isnan(x).all() and self.problem_type == 1 else x,
                    state=int(g("state")), w=float(g("w")), res=float(g("res")), prev=float(g("prev")),
                    alpha=alpha, stuck=int(g("stuck")), retries=int(g("retries")), resets=int(g("resets")),
                    hist=int(g("hist")))

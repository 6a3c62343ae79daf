import sys, os, json
sys.path.insert(0, '/root/repo')
import numpy as np
import adaptive_matrix_solver_b200 as pkg
from adaptive_matrix_solver_b200.workloads import k5_sparse, initial_vectors
eng = pkg.MausEngine(0)
for n in (125_000, 250_000, 500_000):
    A = k5_sparse(n); eng.set_matrix(A)
    for C_ in (8, 16):
        V = initial_vectors(C_, n); eng.upload_vectors(V)
        for _ in range(3): eng.rq(C_=C_)
        eng.profile_reset(True)
        for _ in range(20): eng.rq(C_=C_)
        p = eng.profile_read(); eng.profile_reset(False)
        ms = p["matvec_ms"] / 20
        print(json.dumps(dict(n=n, C=C_, pack8=os.environ.get("MAUS_SPMM_PACK8"), ms_per_matvec=round(ms, 4), launches=p["matvec_launches"])), flush=True)

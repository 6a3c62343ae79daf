import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import adaptive_matrix_solver_b200 as pkg
from adaptive_matrix_solver_b200.workloads import k5_sparse, initial_vectors
eng = pkg.MausEngine(0)
A = k5_sparse(1_000_000); eng.set_matrix(A)
V = initial_vectors(4, 1_000_000); eng.upload_vectors(V)
for _ in range(6): eng.rq(C_=4)

"""Dump the CSR structure of the K5 benchmark matrix for profiles/microbench/spmm_ablate.cu:  python dump_k5_csr.py out.bin"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from adaptive_matrix_solver_b200.workloads import k5_sparse   # noqa: E402

A = k5_sparse(1_000_000).tocsr()
with open(sys.argv[1], "wb") as f:
    np.array([A.shape[0], A.nnz], dtype=np.int64).tofile(f)
    A.indptr.astype(np.int64).tofile(f)
    A.indices.astype(np.int32).tofile(f)
print("rows", A.shape[0], "nnz", A.nnz, "row length min/max", np.diff(A.indptr).min(), np.diff(A.indptr).max(), "sorted", A.has_sorted_indices)

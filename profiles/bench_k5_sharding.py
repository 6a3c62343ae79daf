"""BASELINE config #5 (sparse CSC n = 1M, ~20 nnz/row, GMRES / SpMV path) on G GPUs, both ways:
  (a) matrix REPLICATED, candidates sharded (the default multi-GPU mode, no data-path collective),
  (b) matrix ROW-SHARDED (all-gather of x per SpMM, all-reduce per dot product; NCCL over NVLink).
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
             profiles/bench_k5_sharding.py [n] [candidates]
Prints one JSON line per mode on rank 0.  Same total work in both modes: G x `candidates` linear solves A x = b."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                                    # noqa: E402
import torch.distributed as dist                                # noqa: E402
import adaptive_matrix_solver_b200 as pkg                       # noqa: E402
from adaptive_matrix_solver_b200 import _abi                    # noqa: E402
from adaptive_matrix_solver_b200.rowshard import RowShardedOperator   # noqa: E402
from adaptive_matrix_solver_b200.workloads import k5_sparse     # noqa: E402

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
C_ = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo", rank=rank, world_size=world)


def barrier():
    if world > 1:
        dist.barrier()


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


A = k5_sparse(n)
eng = pkg.MausEngine(local)
psi = np.full(C_, 5e-19)
zero = np.zeros(C_, dtype=complex)

# ---- (a) replicated matrix, this rank's own C_ candidates ----------------------------------------------------------
rng = np.random.default_rng(100 + rank)
RHS = rng.random((C_, n)) + 1j * rng.random((C_, n)); RHS /= np.linalg.norm(RHS, axis=1, keepdims=True)
eng.set_matrix(A)
eng.solve_shifted(zero, psi, rng_key=None, method=_abi.METHOD_GMRES, RHS=RHS)             # warm-up
barrier(); t0 = time.perf_counter()
X, st, it = eng.solve_shifted(zero, psi, rng_key=None, method=_abi.METHOD_GMRES, RHS=RHS)
dt_a = max_over_ranks(time.perf_counter() - t0)
rel_a = float(np.linalg.norm(A @ X[0] - RHS[0]) / np.linalg.norm(RHS[0]))
if rank == 0:
    print(json.dumps(dict(mode="replicated matrix, candidates sharded", n=n, n_gpus=world, candidates_total=C_ * world,
                          seconds=round(dt_a, 4), candidate_solves_per_s=round(C_ * world / dt_a, 1), inner_iters=it.tolist(),
                          status=st.tolist(), rel_residual_c0=rel_a)), flush=True)

# ---- (b) row-sharded matrix, all G x C_ candidates solved jointly ----------------------------------------------------
op = RowShardedOperator(eng, rank, world)
op.set_matrix(A)
Call = C_ * world
rng = np.random.default_rng(100)
RHSall = rng.random((Call, n)) + 1j * rng.random((Call, n)); RHSall /= np.linalg.norm(RHSall, axis=1, keepdims=True)
Rloc = op.local(RHSall)
psi_all = np.full(Call, 5e-19); zero_all = np.zeros(Call, dtype=complex)
op.gmres(zero_all, psi_all, Rloc)                                                      # warm-up
barrier(); t0 = time.perf_counter()
Xl, st, it = op.gmres(zero_all, psi_all, Rloc)
dt_b = max_over_ranks(time.perf_counter() - t0)
# residual of candidate 0 from the gathered solution
if world > 1:
    parts = [None] * world
    dist.all_gather_object(parts, Xl[0])
    x0 = np.concatenate(parts)
else:
    x0 = Xl[0]
rel_b = float(np.linalg.norm(A @ x0 - RHSall[0]) / np.linalg.norm(RHSall[0]))
# one SpMM of all candidates alone (all-gather + local SpMM), timed by the library's events
eng.profile_reset(True)
for _ in range(5):
    op.matvec(Rloc)
p = eng.profile_read(); eng.profile_reset(False)
if rank == 0:
    print(json.dumps(dict(mode=op.mode_description(), n=n, n_gpus=world,
                          candidates_total=Call, seconds=round(dt_b, 4), candidate_solves_per_s=round(Call / dt_b, 1),
                          inner_iters=it.tolist()[:8], status=st.tolist()[:8], rel_residual_c0=rel_b,
                          spmm_ms_all_candidates=round(p["matvec_ms"] / max(1, p["matvec_launches"]), 3),
                          replicated_over_rowsharded=round(dt_b / dt_a, 2))), flush=True)
barrier()
eng.close()
if world > 1:
    dist.destroy_process_group()

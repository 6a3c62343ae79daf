"""Row-sharded sparse operator (BASELINE config #5 as worded): host slicing logic on CPU, the world = 1
path on one GPU (bit-identical to the replicated operator), and a 2-rank NCCL run when the box has two GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def crand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def scipy_gmres(H, b):
    counter = []
    x, info = spla.gmres(H, b, x0=b, rtol=1e-8, maxiter=50, callback=lambda r: counter.append(r), callback_type="pr_norm")
    return x, info, len(counter)


# ---- host logic (CPU) ----------------------------------------------------------------------------------------------
def test_row_blocks_tile_the_matrix():
    from adaptive_matrix_solver_b200.rowshard import csr_row_block, row_block
    from adaptive_matrix_solver_b200.workloads import k5_sparse
    n, world = 96, 4
    A = k5_sparse(n, seed=1)
    dense = A.toarray()
    rows = []
    for r in range(world):
        n_, row0, nloc, rowptr, colidx, vals = csr_row_block(A, r, world)
        assert (n_, row0, nloc) == (n, r * 24, 24) == (n, *row_block(n, r, world))
        assert rowptr[0] == 0 and rowptr[-1] == len(colidx) == len(vals)
        blk = sp.csr_matrix((vals, colidx, rowptr), shape=(nloc, n)).toarray()
        rows.append(blk)
    assert np.array_equal(np.vstack(rows), dense)


def test_row_block_rejects_ragged_split():
    from adaptive_matrix_solver_b200.rowshard import row_block
    with pytest.raises(ValueError):
        row_block(10, 0, 4)


# ---- one GPU, world = 1 -----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng():
    import adaptive_matrix_solver_b200 as pkg
    e = pkg.MausEngine(0)
    yield e
    e.close()


@pytest.mark.gpu
def test_world1_matches_replicated_operator_bitwise(eng):
    from adaptive_matrix_solver_b200 import _abi
    from adaptive_matrix_solver_b200.rowshard import RowShardedOperator
    from adaptive_matrix_solver_b200.workloads import k5_sparse
    n, C = 3000, 6
    A = k5_sparse(n, seed=5)
    rng = np.random.default_rng(3)
    RHS = crand(rng, C, n)
    sigma = 0.2 * crand(rng, C)
    psi = np.full(C, 5e-19)
    op = RowShardedOperator(eng, 0, 1)
    op.set_matrix(A)
    Y = op.matvec(RHS)
    Yref = (A @ RHS.T).T
    assert np.abs(Y - Yref).max() <= 1e-13 * np.abs(Yref).max()
    X, st, it = op.gmres(sigma, psi, RHS)
    eng.set_matrix(A)
    X2, st2, it2 = eng.solve_shifted(sigma, psi, rng_key=None, method=_abi.METHOD_GMRES, RHS=RHS)
    assert np.array_equal(st, st2) and np.array_equal(it, it2)
    assert np.array_equal(X, X2)                      # same kernels, same reduction order
    for c in range(C):
        H = sp.csc_matrix(A - sigma[c] * sp.eye(n, format="csc") + psi[c] * sp.identity(n, format="csc"))
        xr, info, nit = scipy_gmres(H, RHS[c])
        assert st[c] == 0 and info == 0 and it[c] == nit
        assert np.linalg.norm(X[c] - xr) <= 1e-10 * np.linalg.norm(xr)


@pytest.mark.gpu
def test_world1_jacobi_and_nonconvergence(eng):
    from adaptive_matrix_solver_b200 import _abi
    from adaptive_matrix_solver_b200.rowshard import RowShardedOperator
    n = 400
    A = sp.csr_matrix(np.roll(np.eye(n), 1, axis=1).astype(np.complex128))
    b = np.zeros((1, n), dtype=np.complex128); b[0, 0] = 1.0
    op = RowShardedOperator(eng, 0, 1)
    op.set_matrix(A)
    X, st, it = op.gmres([0j], [0.0], b)
    xr, info, nit = scipy_gmres(A.tocsc(), b[0])
    assert info != 0 and st[0] == _abi.ST_GMRES_NOCONV and it[0] == nit
    # Jacobi: diagonal spread over 4 decades
    rng = np.random.default_rng(8)
    n = 512
    d = np.logspace(0, 4, n) * np.exp(1j * rng.uniform(0, 0.3, n))
    A = sp.diags(d) + 0.05 * sp.random(n, n, density=0.02, random_state=3, dtype=np.float64)
    A = sp.csr_matrix(A, dtype=np.complex128)
    RHS = crand(rng, 2, n)
    op.set_matrix(A)
    X, st, it = op.gmres([0j, 0j], [1e-19, 1e-19], RHS, use_jacobi=[1, 0])
    H = sp.csc_matrix(A + 1e-19 * sp.identity(n))
    M = sp.diags(1.0 / H.diagonal())
    counter = []
    xr, info = spla.gmres(H, RHS[0], x0=RHS[0], rtol=1e-8, maxiter=50, M=M, callback=lambda r: counter.append(r),
                          callback_type="pr_norm")
    assert info == 0 and st[0] == 0 and it[0] == len(counter)
    assert np.linalg.norm(X[0] - xr) <= 1e-9 * np.linalg.norm(xr)
    assert it[0] < it[1] or st[1] != 0


# ---- two GPUs, NCCL ----------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _nccl_worker(rank, world, port, out_dir, n, C):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # only carries the 128-byte NCCL id
    import adaptive_matrix_solver_b200 as pkg
    from adaptive_matrix_solver_b200.rowshard import RowShardedOperator
    from adaptive_matrix_solver_b200.workloads import k5_sparse
    A = k5_sparse(n, seed=5)
    rng = np.random.default_rng(3)
    RHS = crand(rng, C, n)
    sigma = 0.2 * crand(rng, C)
    psi = np.full(C, 5e-19)
    eng_ = pkg.MausEngine(rank)
    op = RowShardedOperator(eng_, rank, world)
    op.set_matrix(A)
    Y = op.matvec(op.local(RHS))
    X, st, it = op.gmres(sigma, psi, op.local(RHS))
    np.savez(os.path.join(out_dir, f"rs{rank}.npz"), Y=Y, X=X, st=st, it=it)
    dist.barrier()
    eng_.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_two_rank_nccl_rowshard_matches_scipy(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from adaptive_matrix_solver_b200.workloads import k5_sparse
    n, C, world = 3000, 6, 2
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path), n, C), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rs{r}.npz") for r in range(world)]
    A = k5_sparse(n, seed=5)
    rng = np.random.default_rng(3)
    RHS = crand(rng, C, n)
    sigma = 0.2 * crand(rng, C)
    psi = np.full(C, 5e-19)
    Y = np.concatenate([p["Y"] for p in parts], axis=1)
    X = np.concatenate([p["X"] for p in parts], axis=1)
    Yref = (A @ RHS.T).T
    assert np.abs(Y - Yref).max() <= 1e-13 * np.abs(Yref).max()
    assert np.array_equal(parts[0]["st"], parts[1]["st"]) and np.array_equal(parts[0]["it"], parts[1]["it"])
    for c in range(C):
        H = sp.csc_matrix(A - sigma[c] * sp.eye(n, format="csc") + psi[c] * sp.identity(n, format="csc"))
        xr, info, nit = scipy_gmres(H, RHS[c])
        assert parts[0]["st"][c] == 0 and info == 0
        assert abs(int(parts[0]["it"][c]) - nit) <= 1        # the cross-rank sum changes the last bits of the dots
        assert np.linalg.norm(X[c] - xr) <= 1e-7 * np.linalg.norm(xr)
        assert np.linalg.norm(H @ X[c] - RHS[c]) <= 1e-8 * np.linalg.norm(RHS[c]) * (1 + 1e-6)

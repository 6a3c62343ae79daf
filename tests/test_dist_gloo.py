"""world_size-2 gloo tests (CPU) of the multi-GPU host path: round-robin candidate sharding, the per-generation
all-gather, replica consistency and the best-eigenpair exchange."""
import os
import random
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir, n, C, gens):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adaptive_matrix_solver_b200.dist import Shard, step_population_sharded, gather_energy_and_best
    from adaptive_matrix_solver_b200.workloads import k2_matrix
    from fake_engine import FakeEngine
    from mock_candidate import MockCandidate, ProblemType
    A = k2_matrix(n, seed=3)
    np.random.seed(1); random.seed(1)                    # every rank builds the same replica of the population
    MockCandidate._next_id = 0
    cands = [MockCandidate(A, ProblemType.EIGENVALUE, n) for _ in range(C)]
    cands[2].state = MockCandidate.State.RETIRED          # not live: must be skipped, shifts the round-robin
    strat = dict(overall_psi_aggression_factor=1.0, max_psi_retries=25, current_convergence_threshold=1e-10)
    know = dict(local_solver_preference="direct_solve", is_sparse_problem=False, is_hermitian=False)
    shard = Shard(rank, world, None)
    eng = FakeEngine()
    for g in range(gens):
        live = step_population_sharded(cands, A, None, strat, know, eng, shard)
    resid = np.array([c.residual_k for c in cands if c.state != MockCandidate.State.RETIRED])
    lam = np.array([c.lambda_k for c in cands if c.state != MockCandidate.State.RETIRED])
    V = np.stack([c.v_k for c in cands if c.state != MockCandidate.State.RETIRED])
    mine = shard.owned(len(resid))
    ar, al, best_vec, best_rank = gather_energy_and_best(shard, resid[mine], lam[mine], V[mine])
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), resid=resid, lam=lam, V=V, alpha=np.array([complex(c.alpha_local_step) for c in cands]),
             stuck=np.array([c.stuck_counter for c in cands]), hist=np.array([len(c.residual_history) for c in cands]),
             state=np.array([c.state.value for c in cands]), calls=eng.calls, live=live, ar=ar, al=al, best_vec=best_vec,
             best_rank=best_rank, n_mine=len(mine))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_population_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    n, C, gens = 24, 7, 3
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), n, C, gens), nprocs=2, join=True)
    r0 = np.load(tmp_path / "r0.npz"); r1 = np.load(tmp_path / "r1.npz")
    # replicas agree bit for bit after every generation's all-gather
    for k in ("resid", "lam", "V", "alpha", "stuck", "hist", "state"):
        assert np.array_equal(r0[k], r1[k]), k
    assert int(r0["live"]) == C - 1
    assert int(r0["n_mine"]) + int(r1["n_mine"]) == C - 1 and abs(int(r0["n_mine"]) - int(r1["n_mine"])) <= 1
    assert np.array_equal(r0["hist"][[0, 1, 3]], np.full(3, 1 + gens))      # one history entry per generation, no double append
    # single-process reference run of the same population
    sys.path.insert(0, HERE)
    from adaptive_matrix_solver_b200.dist import Shard, step_population_sharded
    from adaptive_matrix_solver_b200.workloads import k2_matrix
    from fake_engine import FakeEngine
    from mock_candidate import MockCandidate, ProblemType
    A = k2_matrix(n, seed=3)
    np.random.seed(1); random.seed(1)
    MockCandidate._next_id = 0
    cands = [MockCandidate(A, ProblemType.EIGENVALUE, n) for _ in range(C)]
    cands[2].state = MockCandidate.State.RETIRED
    strat = dict(overall_psi_aggression_factor=1.0, max_psi_retries=25, current_convergence_threshold=1e-10)
    know = dict(local_solver_preference="direct_solve", is_sparse_problem=False, is_hermitian=False)
    eng = FakeEngine()
    for g in range(gens):
        step_population_sharded(cands, A, None, strat, know, eng, Shard(0, 1))
    resid = np.array([c.residual_k for c in cands if c.state != MockCandidate.State.RETIRED])
    V = np.stack([c.v_k for c in cands if c.state != MockCandidate.State.RETIRED])
    assert np.array_equal(resid, r0["resid"]) and np.array_equal(V, r0["V"])
    # the energy gather returns every candidate's residual, and the best eigenpair comes from the right rank
    assert sorted(r0["ar"].tolist()) == sorted(resid.tolist())
    assert np.array_equal(r0["ar"], r1["ar"]) and np.array_equal(r0["best_vec"], r1["best_vec"])
    k = int(np.argmin(resid))
    assert np.array_equal(r0["best_vec"], V[k])
    assert int(r0["best_rank"]) == k % 2


def _worker_fail(rank, world, port, out_dir, n, C, gens):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adaptive_matrix_solver_b200.dist import Shard, step_population_sharded
    from adaptive_matrix_solver_b200.workloads import k2_matrix
    from fake_engine import FakeEngine
    from mock_candidate import MockCandidate, ProblemType
    A = k2_matrix(n, seed=3)
    np.random.seed(1); random.seed(1)
    MockCandidate._next_id = 0
    cands = [MockCandidate(A, ProblemType.EIGENVALUE, n) for _ in range(C)]
    strat = dict(overall_psi_aggression_factor=1.0, max_psi_retries=3, current_convergence_threshold=1e-10)
    know = dict(local_solver_preference="direct_solve", is_sparse_problem=False, is_hermitian=False)
    shard = Shard(rank, world, None)
    eng = FakeEngine()
    eng.fail_ids = {1}                                    # owned by rank 1: its ladder is exhausted -> re-initialised from rank 1's RNG
    draws = []
    for g in range(gens):
        step_population_sharded(cands, A, None, strat, know, eng, shard)
        # what the reference's _manage_candidates would draw next on this replica (AMS:525-549)
        draws.append([np.random.rand(), random.random()])
    np.savez(os.path.join(out_dir, f"f{rank}.npz"), draws=np.array(draws), V=np.stack([c.v_k for c in cands]),
             lam=np.array([c.lambda_k for c in cands]), stuck=np.array([c.stuck_counter for c in cands]),
             state=np.array([c.state.value for c in cands]), resets=np.array([c.num_resets for c in cands]))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_replicas_stay_identical_after_a_forced_failure(tmp_path):
    """A candidate whose solves all fail walks the ladder and is re-initialised from the OWNING rank's host RNG (AMS:287-293).
    The replicas must still agree afterwards -- population AND the global RNG streams the next spawn would use."""
    import torch.multiprocessing as mp
    n, C, gens = 16, 5, 2
    port = _free_port()
    mp.spawn(_worker_fail, args=(2, port, str(tmp_path), n, C, gens), nprocs=2, join=True)
    r0 = np.load(tmp_path / "f0.npz"); r1 = np.load(tmp_path / "f1.npz")
    for k in ("V", "lam", "stuck", "state", "resets", "draws"):
        assert np.array_equal(r0[k], r1[k]), k
    assert r0["stuck"][1] == gens and r0["state"][1] == 3          # STUCK, counted once per generation

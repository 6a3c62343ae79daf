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


class _RecycledGather:
    """gloo stand-in of RowShardedOperator.gather: with ``pinned=True`` the result lands in one of two buffers that are reused
    alternately -- exactly the ownership rule the sharded step has to live with on the GPU path"""

    def __init__(self, world):
        self.world = world
        self._bufs = None
        self._turn = 0
        self.calls = 0

    def gather(self, send, pinned=False):
        import torch
        import torch.distributed as dist
        send = np.ascontiguousarray(send, dtype=np.float64).ravel()
        parts = [torch.empty(send.size, dtype=torch.float64) for _ in range(self.world)]
        dist.all_gather(parts, torch.from_numpy(send.copy()))
        if not pinned:
            return np.stack([p.numpy() for p in parts])
        if self._bufs is None or self._bufs[0].shape[1] < send.size:
            self._bufs = [np.full((self.world, send.size), np.nan) for _ in range(2)]
        out = self._bufs[self._turn][:, :send.size]
        self._turn ^= 1
        self.calls += 1
        for r in range(self.world):
            out[r] = parts[r].numpy()
        return out


def _worker_views(rank, world, port, out_dir, n, C, gens):
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
    strat = dict(overall_psi_aggression_factor=1.0, max_psi_retries=25, current_convergence_threshold=1e-10)
    know = dict(local_solver_preference="direct_solve", is_sparse_problem=False, is_hermitian=False)
    eng = FakeEngine()
    eng.rowshard = _RecycledGather(world)
    eng._close_hooks = []
    shard = Shard(rank, world, None, engine=eng)
    step_population_sharded(cands, A, None, strat, know, eng, shard)
    step_population_sharded(cands, A, None, strat, know, eng, shard)
    # "host logic" (the reference's _manage_candidates) retires candidate 3 on every replica between two generations; on the
    # rank that does not own it, its vector is a view into a recycled gather buffer at this moment
    victim = cands[3]
    victim.state = MockCandidate.State.RETIRED
    frozen = victim.v_k.copy()
    was_view = victim.v_k.base is not None
    for g in range(gens):
        step_population_sharded(cands, A, None, strat, know, eng, shard)
    kept = bool(np.array_equal(victim.v_k, frozen))
    own_after = victim.v_k.base is None
    views_before_close = sum(1 for c in cands if c.v_k.base is not None)
    snapshot = np.stack([c.v_k.copy() for c in cands])
    for hook in eng._close_hooks:
        hook()
    for b in eng.rowshard._bufs:
        b[:] = np.nan                                        # "the engine released its page-locked memory"
    intact = bool(np.array_equal(np.stack([c.v_k for c in cands]), snapshot))
    np.savez(os.path.join(out_dir, f"v{rank}.npz"), V=snapshot, kept=kept, own_after=own_after, was_view=was_view,
             views_before_close=views_before_close, intact=intact, gathers=eng.rowshard.calls,
             views_after=sum(1 for c in cands if c.v_k.base is not None))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_recycled_gather_buffers_never_overwrite_a_kept_vector(tmp_path):
    """The GPU path gathers into two recycled page-locked buffers and hands out views: a candidate that leaves the live set through
    host logic must keep its vector over later generations, and every view must be detached before the engine frees the buffers."""
    import torch.multiprocessing as mp
    n, C, gens = 24, 7, 4
    port = _free_port()
    mp.spawn(_worker_views, args=(2, port, str(tmp_path), n, C, gens), nprocs=2, join=True)
    r0 = np.load(tmp_path / "v0.npz"); r1 = np.load(tmp_path / "v1.npz")
    assert np.array_equal(r0["V"], r1["V"])                  # replicas agree
    assert bool(r0["was_view"]) != bool(r1["was_view"])      # exactly one rank held the victim as a view
    for r in (r0, r1):
        assert bool(r["kept"]) and bool(r["own_after"]) and bool(r["intact"])
        assert int(r["gathers"]) == 2 + gens and int(r["views_before_close"]) > 0 and int(r["views_after"]) == 0
